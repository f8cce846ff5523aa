/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * The reference's Airspy front-end (air.c: float32 REAL samples at IF = Fs/4, mixer x boxcar with
 * partial sums carried across transfers of arbitrary length, air.c:291-341) compiled in place
 * (`#include "air.c"`), with -DWITH_AIR on every TU (channel_t gains a field, acarsdec.h:66-68),
 * a stub libairspy whose only sample rate comes from ref_air_open(), and the same outputmsg sink
 * as ref_harness.c.  Output: oracle/_ref/libacarsref_air_O2.so.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "air.c" /* the reference, in place (needs -DWITH_AIR, -Istub) */

channel_t channel[MAXNBCHANNELS];
unsigned int nbch;
int verbose = 0;
int signalExit = 0;
int gain = 18;

/* ---- stub libairspy ---- */
struct airspy_device { int dummy; };
static struct airspy_device the_dev;
static uint32_t stub_rate = 2500000, stub_freq;
int airspy_list_devices(uint64_t *s, int n) { if (s && n > 0) s[0] = 1; return 1; }
int airspy_open_sn(struct airspy_device **d, uint64_t sn) { (void)sn; *d = &the_dev; return AIRSPY_SUCCESS; }
int airspy_open(struct airspy_device **d) { *d = &the_dev; return AIRSPY_SUCCESS; }
int airspy_close(struct airspy_device *d) { (void)d; return 0; }
int airspy_exit(void) { return 0; }
int airspy_set_sample_type(struct airspy_device *d, enum airspy_sample_type t) { (void)d; (void)t; return 0; }
int airspy_get_samplerates(struct airspy_device *d, uint32_t *b, const uint32_t len)
{ (void)d; if (len == 0) *b = 1; else b[0] = stub_rate; return 0; }
int airspy_set_samplerate(struct airspy_device *d, uint32_t r) { (void)d; (void)r; return 0; }
int airspy_set_packing(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return 0; }
int airspy_set_linearity_gain(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return 0; }
int airspy_set_freq(struct airspy_device *d, const uint32_t f) { (void)d; stub_freq = f; return 0; }
int airspy_r820t_write(struct airspy_device *d, uint8_t r, uint8_t v) { (void)d; (void)r; (void)v; return 0; }
int airspy_start_rx(struct airspy_device *d, airspy_sample_block_cb_fn cb, void *c) { (void)d; (void)cb; (void)c; return -1; }
int airspy_stop_rx(struct airspy_device *d) { (void)d; return 0; }
int airspy_is_streaming(struct airspy_device *d) { (void)d; return 0; }
const char *airspy_error_name(enum airspy_error e) { (void)e; return "stub"; }

/* ---- message sink (same as ref_harness.c) ---- */
typedef struct { int chn, len, err; float lvl; unsigned char txt[250]; unsigned char crc[2]; } ref_msg_t;
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

/* ---- Airspy path: initAirspy + per-channel initMsk/initAcars like acarsdec.c:445-454 ---- */
static int opened;
int ref_air_open(unsigned rate, int nfreq, const char **freq_mhz)
{
	char *argv[MAXNBCHANNELS + 4];
	int n, r;
	if (opened || nfreq > MAXNBCHANNELS) return -1;
	stub_rate = rate;
	argv[0] = "0";
	for (n = 0; n < nfreq; n++) argv[1 + n] = (char *)freq_mhz[n];
	argv[1 + nfreq] = NULL;
	memset(channel, 0, sizeof(channel));
	device = NULL;
	ind = 0;
	r = initAirspy(argv, 0);
	if (r) return r;
	for (n = 0; n < (int)nbch; n++) {
		channel[n].chn = n;
		if ((r = initMsk(&channel[n]))) return r;
		if ((r = initAcars(&channel[n]))) return r;
	}
	opened = 1;
	return 0;
}
unsigned ref_air_fc(void) { return stub_freq; }
unsigned ref_air_mult(void) { return AIRMULT; }
int ref_nbch(void) { return (int)nbch; }
void ref_get_wf(int ch, float *out)
{
	for (unsigned i = 0; i < AIRMULT; i++) { out[2 * i] = crealf(channel[ch].wf[i]); out[2 * i + 1] = cimagf(channel[ch].wf[i]); }
}
static int last_m;
/* one call of the reference's static rx_callback (air.c:291); returns the number of envelope
 * samples it produced per channel */
int ref_air_transfer(float *samples, int count)
{
	airspy_transfer_t t;
	int bo = AIRMULT - ind;
	memset(&t, 0, sizeof(t));
	t.samples = samples;
	t.sample_count = count;
	last_m = 1 + (count - bo) / (int)AIRMULT;
	rx_callback(&t);
	return last_m;
}
void ref_get_dm(int ch, float *out, int n) { memcpy(out, channel[ch].dm_buffer, n * sizeof(float)); }

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
void ref_close(void)
{
	if (!opened) return;
	ref_flush();
	deinitAcars();
	for (unsigned n = 0; n < nbch; n++) { free(channel[n].inb); free(channel[n].dm_buffer); free(channel[n].blk); free(channel[n].wf); }
	memset(channel, 0, sizeof(channel));
	nbch = 0; opened = 0;
	pthread_mutex_lock(&sink_mtx);
	sink_n = 0;
	pthread_mutex_unlock(&sink_mtx);
}
