/*
 * TEST INFRASTRUCTURE — stub SoapySDR that replays a raw interleaved CS16 capture, so the reference's
 * UNMODIFIED soapy.c (initSoapy/runSoapySample/readThreadEntryPoint) and acarsdec.c main() run end
 * to end on a file: build target oracle/_ref/acarsdec_ref_soapy.  The "device string" after -d is
 * the capture path.  readStream hands out at most numElems samples per call, with a short read
 * every few calls (the reference carries the tap index across reads, soapy.c:232-254); at the end of
 * the file it lingers 300 ms so blk_thread can drain, then returns 0 and the reference stops.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "stub/SoapySDR/Device.h"

struct SoapySDRDevice { FILE *f; unsigned calls; };
struct SoapySDRStream { int dummy; };
static struct SoapySDRDevice the_dev;
static struct SoapySDRStream the_stream;

SoapySDRDevice *SoapySDRDevice_makeStrArgs(const char *args)
{
	the_dev.f = fopen(args, "rb");
	return the_dev.f ? &the_dev : NULL;
}
int SoapySDRDevice_unmake(SoapySDRDevice *d) { if (d->f) fclose(d->f); d->f = NULL; return 0; }
const char *SoapySDRDevice_lastError(void) { return "stub: end of capture"; }
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
	(void)s; (void)flags; (void)timeNs; (void)timeoutUs;
	size_t want = numElems;
	if (++d->calls % 3 == 0 && want > 1000) want = want - 777;          /* a ragged read now and then */
	size_t got = d->f ? fread(buffs[0], 2 * sizeof(int16_t), want, d->f) : 0;
	if (got == 0) usleep(300 * 1000);
	return (int)got;
}
