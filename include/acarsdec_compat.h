/*
 * The reference's own C API for the hot path, as exported by libacarsdec_compat.so.
 *
 * A maintainer links acarsdec.c against libacarsdec_compat.so INSTEAD OF msk.c, acars.c and
 * rtl.c (+ librtlsdr): the symbols below are the ones those files define and acarsdec.c (and
 * the other front-ends) call.  Normally this header is not needed at all — the host includes
 * its own acarsdec.h; build the shim with -DACB_USE_REFERENCE_HEADER -I<acarsdec source dir>
 * and the real header is used.  Without that macro the declarations below mirror the ABI of
 * acarsdec.h v3.7 (channel_t acarsdec.h:59-92, msgblk_t acarsdec.h:48-57), including its
 * dependence on the WITH_* front-end macros, so that the shim can be built where the
 * reference tree is absent.  tests/test_compat.py checks sizeof/offsetof against the real
 * header whenever /root/reference is present.
 */
#ifndef ACARSDEC_COMPAT_H
#define ACARSDEC_COMPAT_H

#ifdef ACB_USE_REFERENCE_HEADER
#include "acarsdec.h"
#else

#include <complex.h>
#include <pthread.h>
#include <sys/time.h>

#define MAXNBCHANNELS 16        /* acarsdec.h:30 */
#define INTRATE 12500           /* acarsdec.h:31 */

typedef struct mskblk_s {       /* acarsdec.h:48-57 */
	struct mskblk_s *prev;
	int chn;
	struct timeval tv;
	int len;
	int err;
	float lvl;
	char txt[250];
	unsigned char crc[2];
} msgblk_t;

typedef struct {                /* acarsdec.h:59-92 */
	int chn;
#if defined(WITH_RTL) || defined(WITH_AIR)
	int Fr;
	float complex *wf;
#endif
#if defined(WITH_AIR)
	float complex D;
#endif
#if defined(WITH_SDRPLAY) || defined(WITH_SOAPY)
	float Fr;
	float complex *oscillator;
	float complex D;
	int counter;
#endif
	float *dm_buffer;
	double MskPhi;
	double MskDf;
	float MskClk;
	double MskLvlSum;
	int MskBitCount;
	unsigned int MskS, idx;
	float complex *inb;
	unsigned char outbits;
	int nbits;
	enum { WSYN, SYN2, SOH1, TXT, CRC1, CRC2, END } Acarsstate;
	msgblk_t *blk;
	pthread_t th;
} channel_t;

/* globals owned by the host program (acarsdec.c:34-57) */
extern channel_t channel[MAXNBCHANNELS];
extern unsigned int nbch;
extern int verbose;
extern int signalExit;
#ifdef WITH_RTL
extern int gain, ppm, rtlMult;
#endif
#ifdef WITH_AIR
extern int gain;
#endif
#ifdef WITH_SOAPY
extern int rateMult, freq, ppm;
extern double gain;
#endif
#ifdef WITH_SDRPLAY
extern int gain, ppm;
#endif

/* callback OUT of the library: every repaired block, from one consumer thread (acars.c:209) */
extern void outputmsg(const msgblk_t *);

#endif /* ACB_USE_REFERENCE_HEADER */

#ifdef __cplusplus
extern "C" {
#endif

/* msk.c:30 / msk.c:67 (acarsdec.h:190-191).  demodMSK consumes ch->dm_buffer[0..len), any len,
 * synchronously; all state travels in *ch, so the result does not depend on the chunking. */
int  initMsk(channel_t *ch);
void demodMSK(channel_t *ch, int len);

/* acars.c:218 / 246 / 378 (acarsdec.h:194-196).  initAcars(&channel[0]) starts the consumer
 * thread that feeds outputmsg(); deinitAcars() stops it — after delivering what is queued
 * (the reference drops queued blocks at shutdown, acars.c:109-112). */
int  initAcars(channel_t *ch);
void decodeAcars(channel_t *ch);
int  deinitAcars(void);

#ifdef WITH_RTL
/* rtl.c:193 / 371 / 406 / 413 (acarsdec.h:160-166).  There is no USB device behind this build:
 * argv[optind] names a raw interleaved-u8 IQ capture (rtl_sdr file format, "-" = stdin) sampled
 * at rtlMult*12500 Hz; the following arguments are the channel frequencies in MHz as usual. */
int initRtl(char **argv, int optind);
int runRtlSample(void);
int runRtlCancel(void);
int runRtlClose(void);
#endif

#ifdef WITH_AIR
/* air.c:66 / 344 (acarsdec.h:168-169).  argv[optind] names a raw float32 capture of REAL samples at
 * IF = rate/4 (AIRSPY_SAMPLE_FLOAT32_REAL); rate from ACARSDEC_B200_AIRRATE (default 2500000). */
int initAirspy(char **argv, int optind);
int runAirspySample(void);
#endif

#ifdef WITH_SOAPY
/* soapy.c:69 / 178 / 263 / 297 (acarsdec.h:172-175).  argv[optind] (the device string after -d)
 * names a raw interleaved CS16 capture at rateMult*12500 Hz. */
int initSoapy(char **argv, int optind);
int soapySetAntenna(const char *antenna);
int runSoapySample(void);
int runSoapyClose(void);
#endif

#ifdef WITH_SDRPLAY
/* sdrplay.c:95 / 238.  The capture (interleaved CS16 at 2 MS/s) is named by ACARSDEC_B200_CAPTURE. */
int initSdrplay(char **argv, int optind);
int runSdrplaySample(void);
#endif

#ifdef __cplusplus
}
#endif
#endif
