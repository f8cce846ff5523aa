/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Compiles the UNMODIFIED reference acars.c in place (`#include "acars.c"`) so that the
 * harness can reach its file-static message queue (acars.c:30-33) and push a sentinel
 * block behind everything the decoder has queued: when blk_thread (acars.c:93-215) hands
 * the sentinel to outputmsg(), every earlier block has been processed (single FIFO
 * consumer).  This replaces the "sleep before deinitAcars" a real run relies on
 * (deinitAcars drops queued blocks, acars.c:109-112).  It also exports the reference's
 * tables so tests can pin the generated ones in the product against them.
 */
#include "acars.c" /* the reference, in place */

void ref_tap_push_sentinel(int chn)
{
	msgblk_t *b = malloc(sizeof(msgblk_t));
	unsigned short crc = 0;
	int i;
	memset(b, 0, sizeof(*b));
	b->chn = chn;
	b->len = 13;
	for (i = 0; i < 12; i++) b->txt[i] = 0x01;  /* odd parity */
	b->txt[12] = (char)ETX;                      /* survives the STX/ETX forcing, odd parity */
	for (i = 0; i < b->len; i++) { update_crc(crc, b->txt[i]); }
	b->crc[0] = crc & 0xff;
	b->crc[1] = crc >> 8;
	pthread_mutex_lock(&blkq_mtx);
	b->prev = NULL;
	if (blkq_s) blkq_s->prev = b;
	blkq_s = b;
	if (blkq_e == NULL) blkq_e = blkq_s;
	pthread_cond_signal(&blkq_wcd);
	pthread_mutex_unlock(&blkq_mtx);
}

/* queue an arbitrary pre-FEC block, exactly where decodeAcars would (acars.c:356-364) */
void ref_tap_push_block(int chn, int len, const unsigned char *txt, const unsigned char *crc)
{
	msgblk_t *b = malloc(sizeof(msgblk_t));
	memset(b, 0, sizeof(*b));
	b->chn = chn;
	b->len = len;
	memcpy(b->txt, txt, len > 250 ? 250 : len);
	b->crc[0] = crc[0];
	b->crc[1] = crc[1];
	pthread_mutex_lock(&blkq_mtx);
	b->prev = NULL;
	if (blkq_s) blkq_s->prev = b;
	blkq_s = b;
	if (blkq_e == NULL) blkq_e = blkq_s;
	pthread_cond_signal(&blkq_wcd);
	pthread_mutex_unlock(&blkq_mtx);
}

int ref_tab_syndrom(unsigned short *out, int max)
{
	int n = (int)(sizeof(syndrom) / sizeof(syndrom[0]));
	if (out) for (int i = 0; i < n && i < max; i++) out[i] = syndrom[i];
	return n;
}
void ref_tab_crc(unsigned short *out) { for (int i = 0; i < 256; i++) out[i] = crc_ccitt_table[i]; }
void ref_tab_numbits(unsigned char *out) { for (int i = 0; i < 256; i++) out[i] = numbits[i]; }
