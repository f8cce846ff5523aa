/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * The reference's SoapySDR front-end (soapy.c: CS16 IQ, D += v*osc/32768.0 with D and the tap index
 * carried across reads, soapy.c:232-254) compiled in place with -DWITH_SOAPY and a stub SoapySDR whose
 * readStream hands out caller-supplied samples in caller-chosen read sizes.
 * Output: oracle/_ref/libacarsref_soapy_O2.so.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "soapy.c" /* the reference, in place (needs -DWITH_SOAPY, -Istub) */

channel_t channel[MAXNBCHANNELS];
unsigned int nbch;
int verbose = 0;
int signalExit = 0;
double gain = -10.0;
int ppm = 0;
int rateMult = 160;
int freq = 0;

struct SoapySDRDevice { int dummy; };
struct SoapySDRStream { int dummy; };
static struct SoapySDRDevice the_dev;
static struct SoapySDRStream the_stream;
static const int16_t *feed;          /* interleaved I,Q */
static size_t feed_n, feed_pos;      /* in complex samples */
static const int *feed_sizes;        /* read sizes (complex samples), cycled */
static int feed_nsizes, feed_i;

SoapySDRDevice *SoapySDRDevice_makeStrArgs(const char *a) { (void)a; return &the_dev; }
int SoapySDRDevice_unmake(SoapySDRDevice *d) { (void)d; return 0; }
const char *SoapySDRDevice_lastError(void) { return "stub"; }
int SoapySDRDevice_setGainMode(SoapySDRDevice *d, int dir, size_t ch, bool a) { (void)d; (void)dir; (void)ch; (void)a; return 0; }
int SoapySDRDevice_setGain(SoapySDRDevice *d, int dir, size_t ch, double v) { (void)d; (void)dir; (void)ch; (void)v; return 0; }
int SoapySDRDevice_setFrequencyCorrection(SoapySDRDevice *d, int dir, size_t ch, double v) { (void)d; (void)dir; (void)ch; (void)v; return 0; }
int SoapySDRDevice_setFrequency(SoapySDRDevice *d, int dir, size_t ch, double f, const void *a) { (void)d; (void)dir; (void)ch; (void)f; (void)a; return 0; }
int SoapySDRDevice_setSampleRate(SoapySDRDevice *d, int dir, size_t ch, double r) { (void)d; (void)dir; (void)ch; (void)r; return 0; }
int SoapySDRDevice_setAntenna(SoapySDRDevice *d, int dir, size_t ch, const char *n) { (void)d; (void)dir; (void)ch; (void)n; return 0; }
SoapySDRStream *SoapySDRDevice_setupStream(SoapySDRDevice *d, int dir, const char *f, const size_t *c, size_t n, const void *a)
{ (void)d; (void)dir; (void)f; (void)c; (void)n; (void)a; return &the_stream; }
int SoapySDRDevice_closeStream(SoapySDRDevice *d, SoapySDRStream *s) { (void)d; (void)s; return 0; }
int SoapySDRDevice_activateStream(SoapySDRDevice *d, SoapySDRStream *s, int f, long long t, size_t n) { (void)d; (void)s; (void)f; (void)t; (void)n; return 0; }
int SoapySDRDevice_deactivateStream(SoapySDRDevice *d, SoapySDRStream *s, int f, long long t) { (void)d; (void)s; (void)f; (void)t; return 0; }
int SoapySDRDevice_readStream(SoapySDRDevice *d, SoapySDRStream *s, void *const *buffs, size_t numElems, int *flags, long long *timeNs, long timeoutUs)
{
	(void)d; (void)s; (void)flags; (void)timeNs; (void)timeoutUs;
	size_t want = numElems;
	if (feed_nsizes) { size_t w = (size_t)feed_sizes[feed_i++ % feed_nsizes]; if (w < want) want = w; }
	if (want > feed_n - feed_pos) want = feed_n - feed_pos;
	if (want == 0) return 0;                       /* end of the capture: the reference stops */
	memcpy(buffs[0], feed + 2 * feed_pos, want * 2 * sizeof(int16_t));
	feed_pos += want;
	return (int)want;
}

#include "ref_cs16_common.h"

static int opened;
int ref_soapy_open(int K, int user_freq, int nfreq, const char **freq_mhz)
{
	char *argv[MAXNBCHANNELS + 4];
	int n, r;
	if (opened || nfreq > MAXNBCHANNELS) return -1;
	rateMult = K;
	freq = user_freq;
	argv[0] = "driver=stub";
	for (n = 0; n < nfreq; n++) argv[1 + n] = (char *)freq_mhz[n];
	argv[1 + nfreq] = NULL;
	memset(channel, 0, sizeof(channel));
	current_index = 0;
	r = initSoapy(argv, 0);
	if (r) return r;
	for (n = 0; n < (int)nbch; n++) {
		channel[n].chn = n;
		if ((r = initMsk(&channel[n]))) return r;
		if ((r = initAcars(&channel[n]))) return r;
	}
	opened = 1;
	return 0;
}
int ref_soapy_fc(void) { return freq; }
/* run the reference's read loop (soapy.c:212-262) over `n` complex samples, reads cut as `sizes` says */
void ref_soapy_feed(const int16_t *iq, size_t n, const int *sizes, int nsizes)
{
	feed = iq; feed_n = n; feed_pos = 0; feed_sizes = sizes; feed_nsizes = nsizes; feed_i = 0;
	signalExit = 0;
	readThreadEntryPoint(NULL);
	signalExit = 0;
}
void ref_close(void)
{
	if (!opened) return;
	ref_flush();
	deinitAcars();
	for (unsigned n = 0; n < nbch; n++) { free(channel[n].inb); free(channel[n].dm_buffer); free(channel[n].blk); free(channel[n].oscillator); }
	memset(channel, 0, sizeof(channel));
	free(soapyInBuf); soapyInBuf = NULL;
	nbch = 0; opened = 0; freq = 0;
	pthread_mutex_lock(&sink_mtx); sink_n = 0; pthread_mutex_unlock(&sink_mtx);
}
