/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Harness around the UNMODIFIED reference sources, compiled where they lie under
 * /root/reference (never copied): this TU does `#include "rtl.c"` so that the
 * file-static channelizer `in_callback` (rtl.c:314-361) and `initRtl` (rtl.c:193-312)
 * are reachable; ref_acars_tap.c does the same for acars.c; msk.c is compiled as is.
 * All TUs are built with -DWITH_RTL because channel_t's layout depends on it
 * (acarsdec.h:62-74).  Output: oracle/_ref/libacarsref*.so (git-ignored, travels to
 * the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * What it provides to the tests (ctypes, see tests/refs.py):
 *   - the globals acarsdec.c would define (acarsdec.c:34-57),
 *   - stub librtlsdr entry points (no device),
 *   - ref_open_rtl / ref_rtl_block        : initRtl + in_callback on caller-supplied u8 IQ,
 *   - ref_open_audio / ref_audio_chunk    : the soundfile.c:58-81 loop without libsndfile,
 *   - an outputmsg() sink that records every msgblk_t the reference emits,
 *   - state snapshots of channel_t for trace comparison.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "rtl.c" /* the reference's RTL front-end, in place (needs -DWITH_RTL, -Istub) */

/* ---- globals normally owned by acarsdec.c (acarsdec.c:34-57) ---- */
channel_t channel[MAXNBCHANNELS];
unsigned int nbch;
int verbose = 0;
int signalExit = 0;
int gain = -100;
int ppm = 0;
int rtlMult = 160;

/* ---- stub librtlsdr ---- */
struct rtlsdr_dev { int dummy; };
static struct rtlsdr_dev the_dev;
static uint32_t stub_center_freq, stub_sample_rate;
uint32_t rtlsdr_get_device_count(void) { return 1; }
const char *rtlsdr_get_device_name(uint32_t i) { (void)i; return "oracle-stub"; }
int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s)
{ (void)i; strcpy(m, "stub"); strcpy(p, "stub"); strcpy(s, "00000001"); return 0; }
int rtlsdr_open(rtlsdr_dev_t **d, uint32_t i) { (void)i; *d = &the_dev; return 0; }
int rtlsdr_close(rtlsdr_dev_t *d) { (void)d; return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *d, int m) { (void)d; (void)m; return 0; }
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *d, int *g) { (void)d; if (g) g[0] = 0; return 1; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *d, int g) { (void)d; (void)g; return 0; }
int rtlsdr_set_freq_correction(rtlsdr_dev_t *d, int p) { (void)d; (void)p; return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t *d, uint32_t f) { (void)d; stub_center_freq = f; return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *d, uint32_t r) { (void)d; stub_sample_rate = r; return 0; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *d) { (void)d; return 0; }
int rtlsdr_read_async(rtlsdr_dev_t *d, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t n, uint32_t l)
{ (void)d; (void)cb; (void)ctx; (void)n; (void)l; return -1; }
int rtlsdr_cancel_async(rtlsdr_dev_t *d) { (void)d; return 0; }

/* ---- message sink: the reference's blk_thread calls this (acars.c:209) ---- */
typedef struct {
	int chn, len, err;
	float lvl;
	unsigned char txt[250];
	unsigned char crc[2];
} ref_msg_t;

#define REF_SENTINEL_CHN 0x7ffe
static pthread_mutex_t sink_mtx = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t sink_cnd = PTHREAD_COND_INITIALIZER;
static ref_msg_t *sink;
static int sink_n, sink_cap, sink_sentinels;

void outputmsg(const msgblk_t *blk)
{
	pthread_mutex_lock(&sink_mtx);
	if (blk->chn == REF_SENTINEL_CHN) {
		sink_sentinels++;
		pthread_cond_broadcast(&sink_cnd);
	} else {
		if (sink_n == sink_cap) {
			sink_cap = sink_cap ? 2 * sink_cap : 256;
			sink = realloc(sink, sink_cap * sizeof(ref_msg_t));
		}
		ref_msg_t *m = &sink[sink_n++];
		m->chn = blk->chn; m->len = blk->len; m->err = blk->err; m->lvl = blk->lvl;
		memcpy(m->txt, blk->txt, 250);
		memcpy(m->crc, blk->crc, 2);
	}
	pthread_mutex_unlock(&sink_mtx);
}

extern void ref_tap_push_sentinel(int chn); /* ref_acars_tap.c */

/* wait until every block the reference has queued so far went through blk_thread */
int ref_flush(void)
{
	pthread_mutex_lock(&sink_mtx);
	int want = sink_sentinels + 1;
	pthread_mutex_unlock(&sink_mtx);
	ref_tap_push_sentinel(REF_SENTINEL_CHN);
	pthread_mutex_lock(&sink_mtx);
	while (sink_sentinels < want)
		pthread_cond_wait(&sink_cnd, &sink_mtx);
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

/* ---- RTL path: initRtl + per-channel initMsk/initAcars like acarsdec.c:445-454 ---- */
static int opened;

int ref_open_rtl(int K, int nfreq, const char **freq_mhz)
{
	char *argv[MAXNBCHANNELS + 4];
	int n, r;
	if (opened || nfreq > MAXNBCHANNELS) return -1;
	rtlMult = K;
	argv[0] = "0";
	for (n = 0; n < nfreq; n++) argv[1 + n] = (char *)freq_mhz[n];
	argv[1 + nfreq] = NULL;
	memset(channel, 0, sizeof(channel));
	r = initRtl(argv, 0);
	if (r) return r;
	for (n = 0; n < (int)nbch; n++) {
		channel[n].chn = n;
		if ((r = initMsk(&channel[n]))) return r;
		if ((r = initAcars(&channel[n]))) return r;
	}
	opened = 1;
	return 0;
}

unsigned ref_center_freq(void) { return stub_center_freq; }
int ref_nbch(void) { return (int)nbch; }
int ref_chan_freq(int ch) { return channel[ch].Fr; }

void ref_get_wf(int ch, float *out)
{
	for (int i = 0; i < rtlMult; i++) {
		out[2 * i] = crealf(channel[ch].wf[i]);
		out[2 * i + 1] = cimagf(channel[ch].wf[i]);
	}
}

/* one call of the reference's static in_callback (rtl.c:314) */
void ref_rtl_block(unsigned char *buf, unsigned nread) { in_callback(buf, nread, NULL); }

void ref_get_dm(int ch, float *out, int n) { memcpy(out, channel[ch].dm_buffer, n * sizeof(float)); }

/* ---- audio path: what soundfile.c:58-81 does, minus libsndfile ---- */
int ref_open_audio(int nch)
{
	int n, r;
	if (opened || nch > MAXNBCHANNELS) return -1;
	memset(channel, 0, sizeof(channel));
	nbch = nch;
	for (n = 0; n < nch; n++) {
		channel[n].chn = n;
		channel[n].dm_buffer = malloc(sizeof(float) * 4096);
		if ((r = initMsk(&channel[n]))) return r;
		if ((r = initAcars(&channel[n]))) return r;
	}
	opened = 2;
	return 0;
}

void ref_audio_chunk(int ch, const float *samples, int len)
{
	memcpy(channel[ch].dm_buffer, samples, len * sizeof(float));
	demodMSK(&channel[ch], len);
}

/* ---- state snapshot (for trace comparison against the restatement / the CUDA path) ---- */
typedef struct {
	double MskPhi, MskDf, MskLvlSum;
	float MskClk;
	int MskBitCount;
	unsigned MskS, idx;
	int nbits, state;
	unsigned char outbits;
	float inb[22];
} ref_state_t;

void ref_state(int ch, ref_state_t *s)
{
	channel_t *c = &channel[ch];
	s->MskPhi = c->MskPhi; s->MskDf = c->MskDf; s->MskLvlSum = c->MskLvlSum;
	s->MskClk = c->MskClk; s->MskBitCount = c->MskBitCount;
	s->MskS = c->MskS; s->idx = c->idx; s->nbits = c->nbits; s->state = (int)c->Acarsstate;
	s->outbits = c->outbits;
	for (int i = 0; i < 11; i++) { s->inb[2 * i] = crealf(c->inb[i]); s->inb[2 * i + 1] = cimagf(c->inb[i]); }
}

void ref_close(void)
{
	if (!opened) return;
	ref_flush();
	deinitAcars();
	for (unsigned n = 0; n < nbch; n++) {
		free(channel[n].inb);
		free(channel[n].dm_buffer);
		free(channel[n].blk);
		if (opened == 1) free(channel[n].wf);
	}
	memset(channel, 0, sizeof(channel));
	nbch = 0;
	opened = 0;
	pthread_mutex_lock(&sink_mtx);
	sink_n = 0;
	pthread_mutex_unlock(&sink_mtx);
}

/* throughput loop for the CPU baseline: nblk in_callback calls over a ring of nbuf blocks */
void ref_rtl_run(unsigned char *bufs, unsigned nread, int nbuf, int nblk)
{
	for (int b = 0; b < nblk; b++)
		in_callback(bufs + (size_t)(b % nbuf) * nread, nread, NULL);
}
