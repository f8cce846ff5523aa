/*
 * TEST INFRASTRUCTURE — stub mirsdrapi that replays a raw interleaved CS16 capture through the
 * reference's UNMODIFIED sdrplay.c (initSdrplay/runSdrplaySample/myStreamCallback) under acarsdec.c
 * main(): build target oracle/_ref/acarsdec_ref_sdrplay.  Capture path from ACARSDEC_STUB_SDRPLAY.
 * mir_sdr_StreamInit starts a thread that feeds the stream callback 1344-sample packets (planar
 * xi/xq, as the API delivers them).  runSdrplaySample never returns in the reference, so at the end
 * of the capture the feeder lingers 300 ms for blk_thread and then ends the process with exit(0).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "stub/mirsdrapi-rsp.h"

static FILE *cap;
static mir_sdr_StreamCallback_t stream_cb;
static void *cb_ctx;

mir_sdr_ErrT mir_sdr_ApiVersion(float *v) { *v = MIR_SDR_API_VERSION; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_GetDevices(mir_sdr_DeviceT *d, unsigned int *n, unsigned int max)
{
	(void)max;
	const char *path = getenv("ACARSDEC_STUB_SDRPLAY");
	cap = path ? fopen(path, "rb") : NULL;
	*n = cap ? 1 : 0;
	d[0].SerNo = "0"; d[0].DevNm = "capture"; d[0].hwVer = 255; d[0].devAvail = 1;
	return mir_sdr_Success;
}
mir_sdr_ErrT mir_sdr_SetDeviceIdx(unsigned int i) { (void)i; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_ReleaseDeviceIdx(void) { return mir_sdr_Success; }

static void *feeder(void *arg)
{
	enum { PKT = 1344 };
	short iq[2 * PKT], xi[PKT], xq[PKT];
	unsigned first = 0;
	size_t n;
	(void)arg;
	while ((n = fread(iq, 2 * sizeof(short), PKT, cap)) > 0) {
		for (size_t i = 0; i < n; i++) { xi[i] = iq[2 * i]; xq[i] = iq[2 * i + 1]; }
		stream_cb(xi, xq, first, 0, 0, 0, (unsigned)n, 0, 0, cb_ctx);
		first += (unsigned)n;
	}
	usleep(300 * 1000);
	exit(0);
}

mir_sdr_ErrT mir_sdr_StreamInit(int *gRdB, double fsMHz, double rfMHz, mir_sdr_Bw_MHzT bw, mir_sdr_If_kHzT ift, int lna, int *gsys,
                                mir_sdr_SetGrModeT mode, int *spp, mir_sdr_StreamCallback_t scb, mir_sdr_GainChangeCallback_t gcb, void *ctx)
{
	pthread_t th;
	(void)gRdB; (void)fsMHz; (void)rfMHz; (void)bw; (void)ift; (void)lna; (void)gsys; (void)mode; (void)gcb;
	*spp = 1344;
	stream_cb = scb; cb_ctx = ctx;
	return pthread_create(&th, NULL, feeder, NULL) ? mir_sdr_Fail : mir_sdr_Success;
}
mir_sdr_ErrT mir_sdr_AgcControl(mir_sdr_AgcControlT e, int sp, int k, unsigned int dms, unsigned int hms, int su, int lna)
{ (void)e; (void)sp; (void)k; (void)dms; (void)hms; (void)su; (void)lna; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetPpm(double p) { (void)p; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetDcMode(int a, int b) { (void)a; (void)b; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetDcTrackTime(int t) { (void)t; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_DCoffsetIQimbalanceControl(unsigned int a, unsigned int b) { (void)a; (void)b; return mir_sdr_Success; }
