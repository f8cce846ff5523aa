/*
 * Test host for libacarsdec_compat.so: the part of acarsdec.c a front-end needs — the globals
 * (acarsdec.c:34-57), the outputmsg() sink and a main loop — so the reference-API symbols can be
 * exercised without the reference tree.  Modes:
 *
 *   wavhost audio <f32 file> <nch> [chunk]   soundfile.c:58-81's loop: interleaved float32 envelope
 *                                            samples, `chunk` frames (default 4096) per demodMSK call
 *   wavhost bytes <file>                     every byte handed to decodeAcars() as ch->outbits
 *                                            (the host-side frame synchroniser entry)
 *   wavhost rtl <K> <iq file> <MHz>...       initRtl/runRtlSample/runRtlClose like acarsdec.c:279,484
 *
 * Every delivered block is printed as: chn len err lvl_bits crc txt(hex)
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "acarsdec_compat.h"

channel_t channel[MAXNBCHANNELS];
unsigned int nbch;
int verbose = 0;
int signalExit = 0;
int gain = -100, ppm = 0, rtlMult = 160;

void outputmsg(const msgblk_t *blk)
{
	union { float f; unsigned u; } l;
	l.f = blk->lvl;
	printf("%d %d %d %08x %02x%02x ", blk->chn, blk->len, blk->err, l.u, blk->crc[0], blk->crc[1]);
	for (int i = 0; i < blk->len; i++) printf("%02x", (unsigned char)blk->txt[i]);
	printf("\n");
	fflush(stdout);
}

static int init_decoders(void)
{
	for (unsigned n = 0; n < nbch; n++) {          /* acarsdec.c:445-454 */
		channel[n].chn = n;
		if (initMsk(&channel[n])) return 1;
		if (initAcars(&channel[n])) return 1;
	}
	return 0;
}

int main(int argc, char **argv)
{
	if (argc >= 4 && !strcmp(argv[1], "audio")) {
		FILE *f = fopen(argv[2], "rb");
		nbch = atoi(argv[3]);
		int chunk = argc > 4 ? atoi(argv[4]) : 4096;
		if (!f || nbch < 1 || nbch > MAXNBCHANNELS) return 2;
		for (unsigned n = 0; n < nbch; n++) channel[n].dm_buffer = malloc(sizeof(float) * chunk);
		if (init_decoders()) { fprintf(stderr, "Unable to init internal decoders\n"); return 1; }
		float *buf = malloc(sizeof(float) * chunk * nbch);
		size_t got;
		while ((got = fread(buf, sizeof(float), (size_t)chunk * nbch, f)) > 0) {
			int len = (int)(got / nbch);
			for (unsigned n = 0; n < nbch; n++) {      /* soundfile.c:71-77 */
				for (int i = 0; i < len; i++) channel[n].dm_buffer[i] = buf[n + i * nbch];
				demodMSK(&channel[n], len);
			}
		}
		deinitAcars();
		return 0;
	}
	if (argc >= 3 && !strcmp(argv[1], "bytes")) {
		FILE *f = fopen(argv[2], "rb");
		int c;
		if (!f) return 2;
		nbch = 1;
		channel[0].chn = 0;
		channel[0].inb = NULL;
		if (initAcars(&channel[0])) return 1;
		channel[0].MskLvlSum = 4.0;
		channel[0].MskBitCount = 1;
		while ((c = fgetc(f)) != EOF) {
			channel[0].outbits = (unsigned char)c;
			channel[0].MskLvlSum += 1.0;               /* so that lvl is finite and deterministic */
			channel[0].MskBitCount += 8;
			decodeAcars(&channel[0]);
		}
		deinitAcars();
		return 0;
	}
	if (argc >= 5 && !strcmp(argv[1], "rtl")) {
		rtlMult = atoi(argv[2]);
		if (initRtl(argv, 3)) { fprintf(stderr, "Unable to init input\n"); return 1; }
		if (init_decoders()) { fprintf(stderr, "Unable to init internal decoders\n"); return 1; }
		runRtlSample();
		int res = runRtlClose();
		deinitAcars();
		return res;
	}
	fprintf(stderr, "usage: wavhost audio|bytes|rtl ...\n");
	return 2;
}
