/* TEST INFRASTRUCTURE (oracle/): the reference's own output.c / label.c / netout.c / fileout.c / cJSON.c, compiled in place
 * and unmodified, behind a small driver that feeds them arbitrary msgblk_t records — the expected bytes for
 * tests/test_outfmt.py (the product's own formatter, acarsdec_b200/csrc/outfmt.c, must print the same).
 *
 *   ref_outfmt <outtype> <net: - n N j> <inmode> <airflt> <emptymsg> <labels or -> <station or ->   < records
 *
 * records: { int32 chn, fr_hz, len, err; float lvl; int64 sec, usec; uint8 txt[256]; } until EOF.
 * stdout: per record "\x1e" then whatever outputmsg() printed, then per UDP datagram it sent "\x1d" + the datagram.
 * This file provides the globals acarsdec.c owns (acarsdec.c:34-51) and nothing else of the program. */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>
#include "acarsdec.h"

channel_t channel[MAXNBCHANNELS];
unsigned int nbch = MAXNBCHANNELS;
char *idstation = NULL;
int inmode = 0, verbose = 0, outtype = OUTTYPE_STD, netout = NETLOG_NONE, airflt = 0, emptymsg = 0, mdly = 600;
int hourly = 0, daily = 0, signalExit = 0, skip_reassembly = 1;

extern void build_label_filter(char *arg);

struct rec { int32_t chn, fr_hz, len, err; float lvl; int64_t sec, usec; unsigned char txt[256]; };

int main(int argc, char **argv)
{
	if (argc < 8) return 2;
	outtype = atoi(argv[1]);
	const char net = argv[2][0];
	inmode = atoi(argv[3]);
	airflt = atoi(argv[4]);
	emptymsg = atoi(argv[5]);
	build_label_filter(strcmp(argv[6], "-") ? argv[6] : NULL);
	idstation = strdup(strcmp(argv[7], "-") ? argv[7] : "");
	int sock = -1;
	char rawaddr[64];
	if (net != '-') {
		netout = net == 'N' ? NETLOG_PLANEPLOTTER : net == 'n' ? NETLOG_NATIVE : NETLOG_JSON;   /* acarsdec.c:356-367 */
		sock = socket(AF_INET, SOCK_DGRAM, 0);
		struct sockaddr_in a = { 0 };
		a.sin_family = AF_INET;
		a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
		socklen_t al = sizeof a;
		if (sock < 0 || bind(sock, (struct sockaddr *)&a, sizeof a) || getsockname(sock, (struct sockaddr *)&a, &al)) return 3;
		snprintf(rawaddr, sizeof rawaddr, "127.0.0.1:%d", ntohs(a.sin_port));
	}
	if (initOutput(NULL, sock >= 0 ? rawaddr : NULL)) return 4;
	struct rec r;
	while (fread(&r, sizeof r, 1, stdin) == 1) {
		msgblk_t blk;
		memset(&blk, 0, sizeof blk);
		blk.chn = r.chn;
		blk.tv.tv_sec = r.sec;
		blk.tv.tv_usec = r.usec;
		blk.len = r.len;
		blk.err = r.err;
		blk.lvl = r.lvl;
		memcpy(blk.txt, r.txt, sizeof blk.txt);
		channel[r.chn % MAXNBCHANNELS].Fr = r.fr_hz;
		fputc(0x1e, stdout);
		outputmsg(&blk);
		fflush(stdout);
		if (sock >= 0) {
			char pkt[4096];
			ssize_t n;
			/* loopback: a datagram the reference sent is queued by the time its write() returned */
			while ((n = recv(sock, pkt, sizeof pkt, MSG_DONTWAIT)) >= 0) {
				fputc(0x1d, stdout);
				fwrite(pkt, 1, (size_t)n, stdout);
			}
			fflush(stdout);
		}
	}
	return 0;
}
