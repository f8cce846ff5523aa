/*
 * TEST INFRASTRUCTURE — stub libairspy that replays a raw float32 (real samples) capture, so the
 * reference's UNMODIFIED air.c (initAirspy/runAirspySample/rx_callback) and acarsdec.c main() run
 * end to end on a file: build target oracle/_ref/acarsdec_ref_air.  Capture path and sample rate
 * come from ACARSDEC_STUB_AIR (file) and ACARSDEC_STUB_AIRRATE (Hz, default 2500000).
 * airspy_start_rx feeds rx_callback 65536-sample transfers from a thread; the last, short transfer
 * is delivered too (air.c handles any length); then it lingers 300 ms so blk_thread can drain.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "stub/libairspy/airspy.h"

struct airspy_device { FILE *f; volatile int streaming; airspy_sample_block_cb_fn cb; pthread_t th; };
static struct airspy_device the_dev;
static uint32_t rate(void) { const char *e = getenv("ACARSDEC_STUB_AIRRATE"); return e ? (uint32_t)atoi(e) : 2500000u; }

int airspy_list_devices(uint64_t *s, int n) { if (s && n > 0) s[0] = 1; return 1; }
int airspy_open_sn(struct airspy_device **d, uint64_t sn)
{
	const char *path = getenv("ACARSDEC_STUB_AIR");
	(void)sn;
	if (!path || !(the_dev.f = fopen(path, "rb"))) { fprintf(stderr, "stub: cannot open ACARSDEC_STUB_AIR\n"); return -1; }
	*d = &the_dev;
	return AIRSPY_SUCCESS;
}
int airspy_open(struct airspy_device **d) { return airspy_open_sn(d, 0); }
int airspy_close(struct airspy_device *d) { if (d->f) fclose(d->f); d->f = NULL; return 0; }
int airspy_exit(void) { return 0; }
int airspy_set_sample_type(struct airspy_device *d, enum airspy_sample_type t) { (void)d; (void)t; return 0; }
int airspy_get_samplerates(struct airspy_device *d, uint32_t *b, const uint32_t len) { (void)d; if (len == 0) *b = 1; else b[0] = rate(); return 0; }
int airspy_set_samplerate(struct airspy_device *d, uint32_t r) { (void)d; (void)r; return 0; }
int airspy_set_packing(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return 0; }
int airspy_set_linearity_gain(struct airspy_device *d, uint8_t v) { (void)d; (void)v; return 0; }
int airspy_set_freq(struct airspy_device *d, const uint32_t f) { (void)d; (void)f; return 0; }
int airspy_r820t_write(struct airspy_device *d, uint8_t r, uint8_t v) { (void)d; (void)r; (void)v; return 0; }
const char *airspy_error_name(enum airspy_error e) { (void)e; return "stub"; }

static void *feeder(void *arg)
{
	struct airspy_device *d = arg;
	float *buf = malloc(65536 * sizeof(float));
	size_t n;
	while ((n = fread(buf, sizeof(float), 65536, d->f)) > 0) {
		airspy_transfer_t t;
		memset(&t, 0, sizeof(t));
		t.device = d; t.samples = buf; t.sample_count = (int)n; t.sample_type = AIRSPY_SAMPLE_FLOAT32_REAL;
		d->cb(&t);
	}
	free(buf);
	usleep(300 * 1000);
	d->streaming = 0;
	return NULL;
}
int airspy_start_rx(struct airspy_device *d, airspy_sample_block_cb_fn cb, void *c)
{
	(void)c;
	d->cb = cb; d->streaming = 1;
	return pthread_create(&d->th, NULL, feeder, d) ? -1 : AIRSPY_SUCCESS;
}
int airspy_stop_rx(struct airspy_device *d) { d->streaming = 0; return 0; }
int airspy_is_streaming(struct airspy_device *d) { return d->streaming ? AIRSPY_TRUE : 0; }
