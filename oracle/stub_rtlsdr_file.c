/*
 * TEST INFRASTRUCTURE — stub librtlsdr that replays a raw u8 IQ capture, so that the reference's
 * UNMODIFIED rtl.c (initRtl/runRtlSample/in_callback) and acarsdec.c main() can run end to end
 * on a file: build target oracle/_ref/acarsdec_ref (see Makefile).  The capture path comes from
 * the environment variable ACARSDEC_STUB_IQ.  rtlsdr_read_async hands rtl.c one buffer of the
 * size it asked for per callback until EOF, then lingers 300 ms so blk_thread can drain its
 * queue (the reference drops queued blocks at shutdown, acars.c:109-112) and returns.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "stub/rtl-sdr.h"

struct rtlsdr_dev { FILE *f; volatile int cancel; };
static struct rtlsdr_dev the_dev;

uint32_t rtlsdr_get_device_count(void) { return 1; }
const char *rtlsdr_get_device_name(uint32_t i) { (void)i; return "file-replay-stub"; }
int rtlsdr_get_device_usb_strings(uint32_t i, char *m, char *p, char *s)
{ (void)i; strcpy(m, "stub"); strcpy(p, "stub"); strcpy(s, "00000001"); return 0; }
int rtlsdr_open(rtlsdr_dev_t **d, uint32_t i)
{
	const char *path = getenv("ACARSDEC_STUB_IQ");
	(void)i;
	if (!path || !(the_dev.f = fopen(path, "rb"))) { fprintf(stderr, "stub: cannot open ACARSDEC_STUB_IQ\n"); return -1; }
	*d = &the_dev;
	return 0;
}
int rtlsdr_close(rtlsdr_dev_t *d) { if (d->f) fclose(d->f); d->f = NULL; return 0; }
int rtlsdr_set_tuner_gain_mode(rtlsdr_dev_t *d, int m) { (void)d; (void)m; return 0; }
int rtlsdr_get_tuner_gains(rtlsdr_dev_t *d, int *g) { (void)d; if (g) g[0] = 0; return 1; }
int rtlsdr_set_tuner_gain(rtlsdr_dev_t *d, int g) { (void)d; (void)g; return 0; }
int rtlsdr_set_freq_correction(rtlsdr_dev_t *d, int p) { (void)d; (void)p; return 0; }
int rtlsdr_set_center_freq(rtlsdr_dev_t *d, uint32_t f) { (void)d; (void)f; return 0; }
int rtlsdr_set_sample_rate(rtlsdr_dev_t *d, uint32_t r) { (void)d; (void)r; return 0; }
int rtlsdr_reset_buffer(rtlsdr_dev_t *d) { (void)d; return 0; }
int rtlsdr_cancel_async(rtlsdr_dev_t *d) { d->cancel = 1; return 0; }
int rtlsdr_read_async(rtlsdr_dev_t *d, rtlsdr_read_async_cb_t cb, void *ctx, uint32_t n, uint32_t len)
{
	unsigned char *buf = malloc(len);
	(void)n;
	while (!d->cancel && fread(buf, 1, len, d->f) == len) cb(buf, len, ctx);
	free(buf);
	usleep(300 * 1000);
	return 0;
}
