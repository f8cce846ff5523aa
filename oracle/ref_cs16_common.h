/* TEST INFRASTRUCTURE — outputmsg sink + state snapshot shared by the soapy/sdrplay harnesses
 * (same code as in ref_harness.c; separate libraries because channel_t differs per WITH_*). */
typedef struct { int chn, len, err; float lvl; unsigned char txt[250]; unsigned char crc[2]; } ref_msg_t;
#define REF_SENTINEL_CHN 0x7ffe
static pthread_mutex_t sink_mtx = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t sink_cnd = PTHREAD_COND_INITIALIZER;
static ref_msg_t *sink;
static int sink_n, sink_cap, sink_sentinels;
void outputmsg(const msgblk_t *blk)
{
	pthread_mutex_lock(&sink_mtx);
	if (blk->chn == REF_SENTINEL_CHN) { sink_sentinels++; pthread_cond_broadcast(&sink_cnd); }
	else {
		if (sink_n == sink_cap) { sink_cap = sink_cap ? 2 * sink_cap : 256; sink = realloc(sink, sink_cap * sizeof(ref_msg_t)); }
		ref_msg_t *m = &sink[sink_n++];
		m->chn = blk->chn; m->len = blk->len; m->err = blk->err; m->lvl = blk->lvl;
		memcpy(m->txt, blk->txt, 250); memcpy(m->crc, blk->crc, 2);
	}
	pthread_mutex_unlock(&sink_mtx);
}
extern void ref_tap_push_sentinel(int chn);
int ref_flush(void)
{
	pthread_mutex_lock(&sink_mtx);
	int want = sink_sentinels + 1;
	pthread_mutex_unlock(&sink_mtx);
	ref_tap_push_sentinel(REF_SENTINEL_CHN);
	pthread_mutex_lock(&sink_mtx);
	while (sink_sentinels < want) pthread_cond_wait(&sink_cnd, &sink_mtx);
	pthread_mutex_unlock(&sink_mtx);
	return 0;
}
int ref_msgs(ref_msg_t *out, int max)
{
	pthread_mutex_lock(&sink_mtx);
	int n = sink_n < max ? sink_n : max;
	if (out) memcpy(out, sink, n * sizeof(ref_msg_t));
	memmove(sink, sink + n, (sink_n - n) * sizeof(ref_msg_t));
	sink_n -= n;
	pthread_mutex_unlock(&sink_mtx);
	return n;
}
typedef struct {
	double MskPhi, MskDf, MskLvlSum; float MskClk; int MskBitCount; unsigned MskS, idx; int nbits, state;
	unsigned char outbits; float inb[22];
} ref_state_t;
void ref_state(int ch, ref_state_t *s)
{
	channel_t *c = &channel[ch];
	s->MskPhi = c->MskPhi; s->MskDf = c->MskDf; s->MskLvlSum = c->MskLvlSum; s->MskClk = c->MskClk;
	s->MskBitCount = c->MskBitCount; s->MskS = c->MskS; s->idx = c->idx; s->nbits = c->nbits;
	s->state = (int)c->Acarsstate; s->outbits = c->outbits;
	for (int i = 0; i < 11; i++) { s->inb[2 * i] = crealf(c->inb[i]); s->inb[2 * i + 1] = cimagf(c->inb[i]); }
}
int ref_nbch(void) { return (int)nbch; }
int ref_counter(int ch) { return channel[ch].counter; }
void ref_get_dm(int ch, float *out, int n) { memcpy(out, channel[ch].dm_buffer, n * sizeof(float)); }
void ref_get_osc(int ch, float *out, int K)
{
	for (int i = 0; i < K; i++) { out[2 * i] = crealf(channel[ch].oscillator[i]); out[2 * i + 1] = cimagf(channel[ch].oscillator[i]); }
}
void ref_get_carry(int ch, float *out) { out[0] = crealf(channel[ch].D); out[1] = cimagf(channel[ch].D); }
