/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 * The reference's SDRplay front-end (sdrplay.c: planar int16 xi/xq, D += v*osc, envelope cabsf(D)/4,
 * D and the tap index carried across callbacks, sdrplay.c:200-236) compiled in place with
 * -DWITH_SDRPLAY and a stub mirsdrapi.  Output: oracle/_ref/libacarsref_sdrplay_O2.so.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "sdrplay.c" /* the reference, in place (needs -DWITH_SDRPLAY, -Istub) */

channel_t channel[MAXNBCHANNELS];
unsigned int nbch;
int verbose = 0;
int signalExit = 0;
int lnaState = 2, GRdB = 20, ppm = 0;

static char stub_name[] = "stub", stub_ser[] = "0001";
mir_sdr_ErrT mir_sdr_ApiVersion(float *v) { *v = MIR_SDR_API_VERSION; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_GetDevices(mir_sdr_DeviceT *d, unsigned int *n, unsigned int max) { (void)max; d[0].SerNo = stub_ser; d[0].DevNm = stub_name; d[0].hwVer = 255; d[0].devAvail = 1; *n = 1; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetDeviceIdx(unsigned int i) { (void)i; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_ReleaseDeviceIdx(void) { return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_StreamInit(int *g, double fs, double rf, mir_sdr_Bw_MHzT bw, mir_sdr_If_kHzT ift, int lna, int *gs, mir_sdr_SetGrModeT m, int *spp,
                                mir_sdr_StreamCallback_t cb, mir_sdr_GainChangeCallback_t gcb, void *ctx)
{ (void)g; (void)fs; (void)rf; (void)bw; (void)ift; (void)lna; (void)gs; (void)m; (void)spp; (void)cb; (void)gcb; (void)ctx; return mir_sdr_Fail; }
mir_sdr_ErrT mir_sdr_AgcControl(mir_sdr_AgcControlT e, int a, int b, unsigned int c, unsigned int d, int f, int g) { (void)e; (void)a; (void)b; (void)c; (void)d; (void)f; (void)g; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetPpm(double p) { (void)p; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetDcMode(int a, int b) { (void)a; (void)b; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_SetDcTrackTime(int t) { (void)t; return mir_sdr_Success; }
mir_sdr_ErrT mir_sdr_DCoffsetIQimbalanceControl(unsigned int a, unsigned int b) { (void)a; (void)b; return mir_sdr_Success; }

#include "ref_cs16_common.h"

static int opened;
int ref_sdrplay_open(int nfreq, const char **freq_mhz)
{
	char *argv[MAXNBCHANNELS + 4];
	int n, r;
	if (opened || nfreq > MAXNBCHANNELS) return -1;
	for (n = 0; n < nfreq; n++) argv[n] = (char *)freq_mhz[n];
	argv[nfreq] = NULL;
	memset(channel, 0, sizeof(channel));
	current_index = 0;
	r = initSdrplay(argv, 0);
	if (r) return r;
	for (n = 0; n < (int)nbch; n++) {
		channel[n].chn = n;
		if ((r = initMsk(&channel[n]))) return r;
		if ((r = initAcars(&channel[n]))) return r;
	}
	opened = 1;
	return 0;
}
unsigned ref_sdrplay_fc(void) { return Fc; }
/* one call of the reference's static stream callback (sdrplay.c:200) */
void ref_sdrplay_packet(int16_t *xi, int16_t *xq, unsigned n) { myStreamCallback(xi, xq, 0, 0, 0, 0, n, 0, 0, NULL); }
void ref_close(void)
{
	if (!opened) return;
	ref_flush();
	deinitAcars();
	for (unsigned n = 0; n < nbch; n++) { free(channel[n].inb); free(channel[n].dm_buffer); free(channel[n].blk); free(channel[n].oscillator); }
	memset(channel, 0, sizeof(channel));
	nbch = 0; opened = 0;
	pthread_mutex_lock(&sink_mtx); sink_n = 0; pthread_mutex_unlock(&sink_mtx);
}
