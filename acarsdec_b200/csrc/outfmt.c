/*
 * Message formatting: SURVEY.md §8 f3.  What the reference does with a repaired block on the way out — outputmsg()'s split
 * of msgblk_t.txt into mode / aircraft / ack / label / block id / message number / flight / text (output.c:486-640), the
 * OOOI fields label.c pulls out of a few labels' texts, and the wire formats: full text (printmsg, output.c:162-224), one
 * line (printoneline, :327-348), JSON (buildjson, :227-324, as cJSON prints it unformatted) and the three UDP payloads of
 * netout.c:101-153 — for hosts that take decoded blocks from libacars_b200 and do NOT link the reference's output.c (the
 * unmodified acarsdec.c host keeps using its own).  Plain C, host only, nothing of it is on the GPU path.
 *
 * Contract: byte-identical to the reference on the same block and options (tests/test_outfmt.py diffs against output.c /
 * label.c / netout.c / cJSON.c compiled in place), with two stated exceptions where the reference's behaviour is undefined:
 *   - label.c reads fixed offsets of the text without looking at its length; here bytes past the end of the text read as 0;
 *   - a block shorter than 13 bytes cannot come out of blk_thread (acars.c:124-129) and is refused.
 * The flight table behind the route JSON (-o 5) and the monitor screen (-o 3) is the acb_flights_* object at the end
 * (output.c:349-484).  Not covered: MQTT, libacars decoding, log-file rotation.
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/acars_b200.h"

/* ---------------------------------------------------------------------------------------- a bounded byte sink */

typedef struct { char *p; size_t cap, n; int overflow; } sink_t;

static void put(sink_t *s, const void *src, size_t len)
{
	if (s->n + len + 1 > s->cap) { s->overflow = 1; return; }
	memcpy(s->p + s->n, src, len);
	s->n += len;
}

static void putc1(sink_t *s, char c) { put(s, &c, 1); }
static void puts0(sink_t *s, const char *z) { put(s, z, strlen(z)); }

/* printf into the sink.  The reference prints through fprintf, which writes what the conversions produce — a %c of NUL
 * included — so the length vsnprintf reports is the length that goes out. */
static void putf(sink_t *s, const char *fmt, ...)
{
	char tmp[1024];
	va_list ap;
	va_start(ap, fmt);
	const int n = vsnprintf(tmp, sizeof tmp, fmt, ap);
	va_end(ap);
	if (n < 0 || (size_t)n >= sizeof tmp) { s->overflow = 1; return; }
	put(s, tmp, (size_t)n);
}

/* ---------------------------------------------------------------------------------------- field split */

static int is_downlink(char bid) { return bid >= '0' && bid <= '9'; }      /* output.c:31 */

/* label.c:24-37 — "H1:Q0:_d" style list; an absent or empty list passes everything */
static int label_passes(const char *list, const char *label)
{
	if (!list || !*list) return 1;
	const size_t ll = strlen(label);
	for (const char *p = list; *p;) {
		const char *e = strchr(p, ':');
		const size_t n = e ? (size_t)(e - p) : strlen(p);
		if (n == ll && memcmp(p, label, n) == 0) return 1;
		if (!e) break;
		p = e + 1;
	}
	return 0;
}

/* The OOOI rules of label.c:39-311 as data: a label, up to five literal guards (offset, text) that must all hold, and up
 * to six four-character fields copied from fixed offsets.  Evaluated in order — a guard that follows a copy in the
 * reference follows it here — which only matters for what a FAILED rule leaves behind, and a failed rule's fields are
 * discarded (DecodeLabel's caller prints nothing then). */
enum { F_DA, F_SA, F_ETA, F_GOUT, F_GIN, F_WOFF, F_WON };
typedef struct { signed char field; short off; } take_t;
typedef struct { short off; const char *lit; } guard_t;
typedef struct {
	char l0, l1;
	guard_t g[5];
	take_t t[6];
} rule_t;
#define END { -1, 0 }
#define NOG { -1, NULL }

static const rule_t RULES[] = {
	/* Q-labels: out / off / on / in reports, label.c:39-166 */
	{ 'Q', '1', { NOG }, { { F_SA, 0 }, { F_GOUT, 4 }, { F_WOFF, 8 }, { F_WON, 12 }, { F_GIN, 16 }, { F_DA, 24 } } },
	{ 'Q', '2', { NOG }, { { F_SA, 0 }, { F_ETA, 4 }, END } },
	{ 'Q', 'A', { NOG }, { { F_SA, 0 }, { F_GOUT, 4 }, END } },
	{ 'Q', 'B', { NOG }, { { F_SA, 0 }, { F_WOFF, 4 }, END } },
	{ 'Q', 'C', { NOG }, { { F_SA, 0 }, { F_WON, 4 }, END } },
	{ 'Q', 'D', { NOG }, { { F_SA, 0 }, { F_GIN, 4 }, END } },
	{ 'Q', 'E', { NOG }, { { F_SA, 0 }, { F_GOUT, 4 }, { F_DA, 8 }, END } },
	{ 'Q', 'F', { NOG }, { { F_SA, 0 }, { F_WOFF, 4 }, { F_DA, 8 }, END } },
	{ 'Q', 'G', { NOG }, { { F_SA, 0 }, { F_GOUT, 4 }, { F_GIN, 8 }, END } },
	{ 'Q', 'H', { NOG }, { { F_SA, 0 }, { F_GOUT, 4 }, END } },
	{ 'Q', 'K', { NOG }, { { F_SA, 0 }, { F_WON, 4 }, { F_DA, 8 }, END } },
	{ 'Q', 'L', { NOG }, { { F_DA, 0 }, { F_GIN, 8 }, { F_SA, 13 }, END } },
	{ 'Q', 'M', { NOG }, { { F_DA, 0 }, { F_SA, 8 }, END } },
	{ 'Q', 'N', { NOG }, { { F_DA, 4 }, { F_ETA, 8 }, END } },
	{ 'Q', 'P', { NOG }, { { F_SA, 0 }, { F_DA, 4 }, { F_GOUT, 8 }, END } },
	{ 'Q', 'Q', { NOG }, { { F_SA, 0 }, { F_DA, 4 }, { F_WOFF, 8 }, END } },
	{ 'Q', 'R', { NOG }, { { F_SA, 0 }, { F_DA, 4 }, { F_WON, 8 }, END } },
	{ 'Q', 'S', { NOG }, { { F_SA, 0 }, { F_DA, 4 }, { F_GIN, 8 }, END } },
	{ 'Q', 'T', { NOG }, { { F_SA, 0 }, { F_DA, 4 }, { F_GOUT, 8 }, { F_GIN, 12 }, END } },
	/* airline-specific free-text layouts, label.c:168-311 (label 26 / RB and 44 have code of their own below) */
	{ '1', '0', { { 0, "ARR01" }, NOG }, { { F_DA, 12 }, { F_ETA, 16 }, END } },
	{ '1', '1', { { 13, "/DS " }, { 21, "/ETA " }, NOG }, { { F_DA, 17 }, { F_ETA, 26 }, END } },
	{ '1', '2', { { 4, "," }, NOG }, { { F_SA, 0 }, { F_DA, 5 }, END } },
	{ '1', '5', { { 0, "FST01" }, NOG }, { { F_SA, 5 }, { F_DA, 9 }, END } },
	{ '1', '7', { { 0, "ETA " }, { 8, "," }, { 13, "," }, NOG }, { { F_ETA, 4 }, { F_SA, 9 }, { F_DA, 14 }, END } },
	{ '1', 'G', { { 4, "," }, NOG }, { { F_SA, 0 }, { F_DA, 5 }, END } },
	{ '2', '0', { { 0, "RST" }, NOG }, { { F_SA, 22 }, { F_DA, 26 }, END } },
	{ '2', '1', { { 6, "," }, { 11, "," }, NOG }, { { F_SA, 7 }, { F_DA, 12 }, END } },
	{ '2', 'N', { { 0, "TKO01" }, { 11, "/" }, NOG }, { { F_SA, 20 }, { F_DA, 24 }, END } },
	{ '2', 'Z', { NOG }, { { F_DA, 0 }, END } },
	{ '3', '3', { { 0, "," }, { 20, "," }, { 25, "," }, NOG }, { { F_SA, 21 }, { F_DA, 26 }, END } },
	{ '3', '9', { { 0, "GTA01" }, { 15, "/" }, NOG }, { { F_SA, 24 }, { F_DA, 28 }, END } },
	{ '4', '5', { { 0, "A" }, NOG }, { { F_DA, 1 }, END } },
	{ '8', '0', { { 6, "/DEST" }, NOG }, { { F_DA, 12 }, END } },        /* label.c:249 compares five of "/DEST/"'s six bytes */
	{ '8', '3', { { 4, "," }, NOG }, { { F_SA, 0 }, { F_DA, 5 }, END } },
	{ '8', 'D', { { 4, "," }, { 35, "," }, { 40, "," }, NOG }, { { F_SA, 36 }, { F_DA, 41 }, END } },
	{ '8', 'E', { { 4, "," }, NOG }, { { F_DA, 0 }, { F_ETA, 5 }, END } },
	{ '8', 'S', { { 4, "," }, NOG }, { { F_DA, 0 }, { F_ETA, 5 }, END } },
};

/* byte i of the text as label.c sees it: the calloc'ed copy of output.c:617 holds txt_len bytes and a NUL */
static char tx(const acb_fields_t *f, const unsigned char *base, int i)
{
	return i >= 0 && i < f->txt_len ? (char)base[f->txt_off + i] : 0;
}

static int guard_holds(const acb_fields_t *f, const unsigned char *base, const guard_t *g)
{
	for (int i = 0; g->lit[i]; i++)
		if (tx(f, base, g->off + i) != g->lit[i]) return 0;
	return 1;
}

static void take4(acb_fields_t *f, const unsigned char *base, int field, int off)
{
	char *dst = field == F_DA ? f->da : field == F_SA ? f->sa : field == F_ETA ? f->eta : field == F_GOUT ? f->gout :
	            field == F_GIN ? f->gin : field == F_WOFF ? f->woff : f->won;
	for (int i = 0; i < 4; i++) dst[i] = tx(f, base, off + i);
}

/* position of character c in the text at or after `from`, C-string semantics (stops at the first NUL); -1 if none */
static int find_ch(const acb_fields_t *f, const unsigned char *base, int from, char c)
{
	for (int i = from; i < f->txt_len && base[f->txt_off + i]; i++)
		if ((char)base[f->txt_off + i] == c) return i;
	return -1;
}

/* label 26 and RB (label.c:186-202): "VER/077", next line "SCH/.../SSSS/DDDD...", optional third line "ETA/hhmm" */
static int rule_26(acb_fields_t *f, const unsigned char *b)
{
	static const guard_t ver = { 0, "VER/077" };
	if (!guard_holds(f, b, &ver)) return 0;
	int p = find_ch(f, b, 0, '\n');
	if (p < 0) return 0;
	p++;
	const guard_t sch = { (short)p, "SCH/" };
	if (!guard_holds(f, b, &sch)) return 0;
	p = find_ch(f, b, p + 4, '/');
	if (p < 0) return 0;
	take4(f, b, F_SA, p + 1);
	take4(f, b, F_DA, p + 6);
	p = find_ch(f, b, p, '\n');
	if (p < 0) return 1;
	p++;
	const guard_t eta = { (short)p, "ETA/" };
	if (!guard_holds(f, b, &eta)) return 0;
	take4(f, b, F_ETA, p + 4);
	return 1;
}

/* label 44 (label.c:221-238): optional "00" prefix, POS0/ETA0 report of kind 2 or 3, comma-separated fixed columns; the
 * ETA is taken twice and the later column wins */
static int rule_44(acb_fields_t *f, const unsigned char *b)
{
	int o = 0;
	if (tx(f, b, 0) == '0') {
		if (tx(f, b, 1) != '0') return 0;
		o = 2;
	}
	const guard_t pos = { (short)o, "POS0" }, eta = { (short)o, "ETA0" };
	if (!guard_holds(f, b, &pos) && !guard_holds(f, b, &eta)) return 0;
	if (tx(f, b, o + 4) != '2' && tx(f, b, o + 4) != '3') return 0;
	static const short commas[] = { 23, 28, 33, 38, 43 };
	for (unsigned i = 0; i < sizeof commas / sizeof commas[0]; i++)
		if (tx(f, b, o + commas[i]) != ',') return 0;
	take4(f, b, F_DA, o + 24);
	take4(f, b, F_ETA, o + 44);
	return 1;
}

static void decode_oooi(acb_fields_t *f, const unsigned char *base)
{
	memset(f->da, 0, sizeof f->da); memset(f->sa, 0, sizeof f->sa); memset(f->eta, 0, sizeof f->eta);
	memset(f->gout, 0, sizeof f->gout); memset(f->gin, 0, sizeof f->gin); memset(f->woff, 0, sizeof f->woff);
	memset(f->won, 0, sizeof f->won);
	f->has_oooi = 0;
	const char l0 = f->label[0], l1 = f->label[1];
	if ((l0 == '2' && l1 == '6') || (l0 == 'R' && l1 == 'B')) { f->has_oooi = rule_26(f, base); return; }
	if (l0 == '4' && l1 == '4') { f->has_oooi = rule_44(f, base); return; }
	for (unsigned r = 0; r < sizeof RULES / sizeof RULES[0]; r++) {
		const rule_t *R = &RULES[r];
		if (R->l0 != l0 || R->l1 != l1) continue;
		for (int i = 0; i < 5 && R->g[i].lit; i++)
			if (!guard_holds(f, base, &R->g[i])) return;
		for (int i = 0; i < 6 && R->t[i].field >= 0; i++) take4(f, base, R->t[i].field, R->t[i].off);
		f->has_oooi = 1;
		return;
	}
}

int acb_msg_fields(const acb_msg_t *m, acb_fields_t *f)
{
	if (!m || !f || m->len < 13 || m->len > ACB_TXTMAX) return 0;       /* blk_thread lets nothing shorter through (acars.c:124-129) */
	const unsigned char *p = m->txt;
	memset(f, 0, sizeof *f);
	f->mode = (char)p[0];
	int j = 0;
	for (int i = 1; i <= 7; i++)                      /* the seven address characters, leading dots dropped (output.c:503-509) */
		if (p[i] != '.') f->addr[j++] = (char)p[i];
	f->ack = p[8] == 0x15 ? '!' : (char)p[8];         /* NAK is not printable */
	f->label[0] = (char)p[9];
	f->label[1] = p[10] == 0x7f ? 'd' : (char)p[10];
	f->bid = (char)p[11];
	f->bs = (char)p[12];
	f->be = (char)p[m->len - 1];
	f->downlink = is_downlink(f->bid);
	int k = 13;
	if (f->bs != 0x03) {
		if (f->downlink) {
			int i;
			for (i = 0; i < 4 && k < m->len - 1; i++, k++) f->no[i] = (char)p[k];
			for (i = 0; i < 6 && k < m->len - 1; i++, k++) f->fid[i] = (char)p[k];
		}
		f->txt_off = k;
		f->txt_len = m->len - k - 1 > 0 ? m->len - k - 1 : 0;
	} else {
		f->txt_off = k;
		f->txt_len = 0;
	}
	decode_oooi(f, p);
	return 1;
}

/* ---------------------------------------------------------------------------------------- printers */

/* the text as a C string: up to its first NUL */
static size_t text_strlen(const acb_fields_t *f, const unsigned char *base)
{
	size_t n = 0;
	while ((int)n < f->txt_len && base[f->txt_off + n]) n++;
	return n;
}

static void put_date(sink_t *s, const acb_fmt_opts_t *o)                  /* printdate + printtime, output.c:138-160 */
{
	if (o->tv_sec + o->tv_usec == 0) return;
	struct tm t;
	const time_t sec = (time_t)o->tv_sec;
	gmtime_r(&sec, &t);
	putf(s, "%02d/%02d/%04d ", t.tm_mday, t.tm_mon + 1, t.tm_year + 1900);
	putf(s, "%02d:%02d:%02d.%03ld", t.tm_hour, t.tm_min, t.tm_sec, (long)(o->tv_usec / 1000));
}

static void put_oooi_text(sink_t *s, const acb_fields_t *f)
{
	if (!f->has_oooi) return;
	puts0(s, "##########################\n");
	if (f->da[0]) putf(s, "Destination Airport : %s\n", f->da);
	if (f->sa[0]) putf(s, "Departure Airport : %s\n", f->sa);
	if (f->eta[0]) putf(s, "Estimation Time of Arrival : %s\n", f->eta);
	if (f->gout[0]) putf(s, "Gate out Time : %s\n", f->gout);
	if (f->gin[0]) putf(s, "Gate in Time : %s\n", f->gin);
	if (f->woff[0]) putf(s, "Wheels off Tme : %s\n", f->woff);          /* sic (output.c:212) */
	if (f->won[0]) putf(s, "Wheels on Time : %s\n", f->won);
}

static void fmt_full(sink_t *s, const acb_msg_t *m, const acb_fields_t *f, const acb_fmt_opts_t *o)
{
	if (o->inmode >= 3)
		putf(s, "\n[#%1d (F:%3.3f L:%+5.1f E:%1d) ", m->chn + 1, (int)o->freq_hz / 1000000.0, m->lvl, m->err);
	else
		putf(s, "\n[#%1d (L:%+5.1f E:%1d) ", m->chn + 1, m->lvl, m->err);
	if (o->inmode != 2) put_date(s, o);
	puts0(s, " --------------------------------\n");
	putf(s, "Mode : %1c ", f->mode);
	putf(s, "Label : %2s ", f->label);
	if (f->bid) {
		putf(s, "Id : %1c ", f->bid);
		if (f->ack == '!') puts0(s, "Nak\n");
		else putf(s, "Ack : %1c\n", f->ack);
		putf(s, "Aircraft reg: %s ", f->addr);
		if (f->downlink) {
			putf(s, "Flight id: %s\n", f->fid);
			putf(s, "No: %4s", f->no);
		}
	}
	putc1(s, '\n');
	const size_t tl = text_strlen(f, m->txt);
	if (tl) { put(s, m->txt + f->txt_off, tl); putc1(s, '\n'); }
	if (f->be == 0x17) puts0(s, "ETB\n");
	put_oooi_text(s, f);
}

static void fmt_oneline(sink_t *s, const acb_msg_t *m, const acb_fields_t *f, const acb_fmt_opts_t *o)
{
	putf(s, "#%1d (L:%+5.1f E:%1d) ", m->chn + 1, m->lvl, m->err);
	if (o->inmode != 2) put_date(s, o);
	putf(s, " %7s %6s %1c %2s %4s ", f->addr, f->fid, f->mode, f->label, f->no);
	size_t tl = text_strlen(f, m->txt);
	if (tl > 59) tl = 59;                               /* strncpy(txt, msg->txt, 59) */
	for (size_t i = 0; i < tl; i++) {
		const char c = (char)m->txt[f->txt_off + i];
		putc1(s, c == '\n' || c == '\r' ? ' ' : c);
	}
	putc1(s, '\n');
}

/* a JSON string the way cJSON's print_string_ptr escapes it (cJSON.c:828-950): the C string up to `len` bytes */
static void json_str(sink_t *s, const char *z, size_t len)
{
	putc1(s, '"');
	for (size_t i = 0; i < len && z[i]; i++) {
		const unsigned char c = (unsigned char)z[i];
		switch (c) {
		case '"': puts0(s, "\\\""); break;
		case '\\': puts0(s, "\\\\"); break;
		case '\b': puts0(s, "\\b"); break;
		case '\f': puts0(s, "\\f"); break;
		case '\n': puts0(s, "\\n"); break;
		case '\r': puts0(s, "\\r"); break;
		case '\t': puts0(s, "\\t"); break;
		default:
			if (c < 32) putf(s, "\\u%04x", c);
			else putc1(s, (char)c);
		}
	}
	putc1(s, '"');
}
static void json_key(sink_t *s, int *first, const char *key)
{
	if (!*first) putc1(s, ',');
	*first = 0;
	json_str(s, key, strlen(key));
	putc1(s, ':');
}
static void json_kv_str(sink_t *s, int *first, const char *key, const char *z, size_t len)
{
	json_key(s, first, key);
	json_str(s, z, len);
}
/* cJSON's print_number (cJSON.c:475-540): 15 significant digits if they read back exactly, else 17 */
static void json_kv_num(sink_t *s, int *first, const char *key, double d)
{
	char b[32];
	double back;
	json_key(s, first, key);
	if (d * 0 != 0) { puts0(s, "null"); return; }
	snprintf(b, sizeof b, "%1.15g", d);
	if (sscanf(b, "%lg", &back) != 1 || back != d) snprintf(b, sizeof b, "%1.17g", d);
	puts0(s, b);
}

static void fmt_json(sink_t *s, const acb_msg_t *m, const acb_fields_t *f, const acb_fmt_opts_t *o)
{
	int first = 1;
	char tmp[8];
	putc1(s, '{');
	json_kv_num(s, &first, "timestamp", (double)o->tv_sec + ((double)o->tv_usec) / 1e6);
	if (o->station_id && o->station_id[0]) json_kv_str(s, &first, "station_id", o->station_id, strlen(o->station_id));
	json_kv_num(s, &first, "channel", m->chn);
	const float freq = (float)((int)o->freq_hz / 1000000.0);        /* float freq = channel[chn].Fr / 1000000.0 (output.c:232) */
	snprintf(tmp, sizeof tmp, "%3.3f", freq);                       /* an 8-byte buffer in the reference too: truncates alike */
	json_key(s, &first, "freq"); puts0(s, tmp);
	snprintf(tmp, sizeof tmp, "%2.1f", m->lvl);
	json_key(s, &first, "level"); puts0(s, tmp);
	json_kv_num(s, &first, "error", m->err);
	json_kv_str(s, &first, "mode", &f->mode, 1);
	json_kv_str(s, &first, "label", f->label, 2);
	if (f->bid) {
		json_kv_str(s, &first, "block_id", &f->bid, 1);
		if (f->ack == '!') { json_key(s, &first, "ack"); puts0(s, "false"); }
		else json_kv_str(s, &first, "ack", &f->ack, 1);
		json_kv_str(s, &first, "tail", f->addr, 7);
		if (f->downlink) {
			json_kv_str(s, &first, "flight", f->fid, 6);
			json_kv_str(s, &first, "msgno", f->no, 4);
		}
	}
	const size_t tl = text_strlen(f, m->txt);
	if (tl) json_kv_str(s, &first, "text", (const char *)m->txt + f->txt_off, tl);
	if (f->be == 0x17) { json_key(s, &first, "end"); puts0(s, "true"); }
	if (f->has_oooi) {
		if (f->sa[0]) json_kv_str(s, &first, "depa", f->sa, 4);
		if (f->da[0]) json_kv_str(s, &first, "dsta", f->da, 4);
		if (f->eta[0]) json_kv_str(s, &first, "eta", f->eta, 4);
		if (f->gout[0]) json_kv_str(s, &first, "gtout", f->gout, 4);
		if (f->gin[0]) json_kv_str(s, &first, "gtin", f->gin, 4);
		if (f->woff[0]) json_kv_str(s, &first, "wloff", f->woff, 4);
		if (f->won[0]) json_kv_str(s, &first, "wlin", f->won, 4);
	}
	json_key(s, &first, "app");
	puts0(s, "{\"name\":\"acarsdec\",\"ver\":\"" ACB_REFERENCE_VERSION "\"}");
	putc1(s, '}');
}

/* netout.c builds its datagram with snprintf and sends strlen() bytes of it: anything after a NUL character (a NUL mode or
 * ack byte printed with %c) never leaves */
static void cut_at_nul(sink_t *s, size_t from)
{
	for (size_t i = from; i < s->n; i++)
		if (s->p[i] == 0) { s->n = i; return; }
}

static void fmt_net_pp(sink_t *s, const acb_msg_t *m, const acb_fields_t *f)           /* Netoutpp, netout.c:101-120 */
{
	const size_t from = s->n;
	putf(s, "AC%1c %7s %1c %2s %1c %4s %6s ", f->mode, f->addr, f->ack, f->label, f->bid ? f->bid : '.', f->no, f->fid);
	const size_t tl = text_strlen(f, m->txt);
	for (size_t i = 0; i < tl; i++) {
		const char c = (char)m->txt[f->txt_off + i];
		putc1(s, c == '\n' || c == '\r' ? ' ' : c);
	}
	cut_at_nul(s, from);
}

static void fmt_net_native(sink_t *s, const acb_msg_t *m, const acb_fields_t *f, const acb_fmt_opts_t *o)   /* Netoutsv, :122-141 */
{
	const size_t from = s->n;
	struct tm t;
	const time_t sec = (time_t)o->tv_sec;
	gmtime_r(&sec, &t);
	putf(s, "%8s %1d %02d/%02d/%04d %02d:%02d:%02d %1d %03d %1c %7s %1c %2s %1c %4s %6s ", o->station_id ? o->station_id : "",
	     m->chn + 1, t.tm_mday, t.tm_mon + 1, t.tm_year + 1900, t.tm_hour, t.tm_min, t.tm_sec, m->err, (int)m->lvl, f->mode, f->addr,
	     f->ack, f->label, f->bid ? f->bid : '.', f->no, f->fid);
	put(s, m->txt + f->txt_off, text_strlen(f, m->txt));
	cut_at_nul(s, from);
}

int acb_format_msg(const acb_msg_t *m, int format, const acb_fmt_opts_t *opt, char *out, size_t cap)
{
	static const acb_fmt_opts_t none = { 0, 0, 0, 0, 0, 0, NULL, NULL };
	if (!m || !out || cap == 0) return ACB_ERR_ARG;
	const acb_fmt_opts_t *o = opt ? opt : &none;
	acb_fields_t f;
	if (!acb_msg_fields(m, &f)) return ACB_ERR_ARG;
	/* the filters of outputmsg (output.c:536-539, 651-652) */
	if (o->airflt && !f.downlink) return 0;
	if (!label_passes(o->labels, f.label)) return 0;
	if (o->emptymsg && text_strlen(&f, m->txt) == 0) return 0;
	sink_t s = { out, cap, 0, 0 };
	switch (format) {
	case ACB_FMT_ONELINE: fmt_oneline(&s, m, &f, o); break;
	case ACB_FMT_FULL: fmt_full(&s, m, &f, o); break;
	case ACB_FMT_JSON: fmt_json(&s, m, &f, o); break;
	case ACB_FMT_NET_PP: fmt_net_pp(&s, m, &f); break;
	case ACB_FMT_NET_NATIVE: fmt_net_native(&s, m, &f, o); break;
	case ACB_FMT_NET_JSON: fmt_json(&s, m, &f, o); putc1(&s, '\n'); break;
	default: return ACB_ERR_ARG;
	}
	if (s.overflow) return ACB_ERR_ARG;
	out[s.n] = 0;
	return (int)s.n;
}

/* ---------------------------------------------------------------------------------------- flight table
 * output.c:349-426: one entry per aircraft address seen on a downlink, most recent first; an entry keeps the flight id of
 * its last message, first / last time, the channels it was heard on, a message count and whatever OOOI fields its messages
 * have yielded so far; entries silent for more than `mdly` seconds are dropped whenever a message is added. */
typedef struct flight {
	struct flight *next;
	char addr[8], fid[7];
	int64_t first_sec, first_usec, last_sec;
	int chmask, count, route_sent;
	char da[5], sa[5], eta[5], gout[5], gin[5], woff[5], won[5];
} flight_t;

struct acb_flights { flight_t *head; int mdly; };

acb_flights_t *acb_flights_new(int mdly_seconds)
{
	acb_flights_t *t = (acb_flights_t *)calloc(1, sizeof *t);
	if (t) t->mdly = mdly_seconds > 0 ? mdly_seconds : 600;           /* acarsdec.c:44 */
	return t;
}

void acb_flights_free(acb_flights_t *t)
{
	if (!t) return;
	for (flight_t *f = t->head; f;) {
		flight_t *n = f->next;
		free(f);
		f = n;
	}
	free(t);
}

static void keep4(char *dst, const char *src) { if (src[0]) memcpy(dst, src, 5); }

/* addFlight: only downlinks with a text part reach it (outflg, output.c:553-575) */
static flight_t *flights_add(acb_flights_t *t, const acb_msg_t *m, const acb_fields_t *f, const acb_fmt_opts_t *o)
{
	flight_t *fl = t->head, *prev = NULL;
	while (fl && strcmp(f->addr, fl->addr) != 0) { prev = fl; fl = fl->next; }
	if (!fl) {
		fl = (flight_t *)calloc(1, sizeof *fl);
		if (!fl) return NULL;
		strncpy(fl->addr, f->addr, sizeof fl->addr);
		fl->first_sec = o->tv_sec;
		fl->first_usec = o->tv_usec;
		fl->next = t->head;                                           /* new entries go to the front */
		t->head = fl;
	} else if (prev) {                                                /* a known one moves to the front */
		prev->next = fl->next;
		fl->next = t->head;
		t->head = fl;
	}
	strncpy(fl->fid, f->fid, sizeof fl->fid);
	fl->last_sec = o->tv_sec;
	if (m->chn >= 0 && m->chn < 31) fl->chmask |= 1 << m->chn;
	fl->count++;
	if (f->has_oooi) {
		keep4(fl->da, f->da); keep4(fl->sa, f->sa); keep4(fl->eta, f->eta); keep4(fl->gout, f->gout);
		keep4(fl->gin, f->gin); keep4(fl->woff, f->woff); keep4(fl->won, f->won);
	}
	for (flight_t **pp = &t->head; *pp;) {                            /* expiry */
		if ((*pp)->last_sec < o->tv_sec - t->mdly) {
			flight_t *dead = *pp;
			*pp = dead->next;
			free(dead);
		} else {
			pp = &(*pp)->next;
		}
	}
	return fl;
}

static int flights_gate(const acb_msg_t *m, const acb_fmt_opts_t *o, acb_fields_t *f)
{
	if (!acb_msg_fields(m, f)) return ACB_ERR_ARG;
	if (o->airflt && !f->downlink) return 0;
	if (!label_passes(o->labels, f->label)) return 0;
	return 1;
}

int acb_flights_route_json(acb_flights_t *t, const acb_msg_t *m, const acb_fmt_opts_t *opt, char *out, size_t cap)
{
	static const acb_fmt_opts_t none = { 0, 0, 0, 0, 0, 0, NULL, NULL };
	if (!t || !m || !out || cap == 0) return ACB_ERR_ARG;
	const acb_fmt_opts_t *o = opt ? opt : &none;
	acb_fields_t f;
	const int g = flights_gate(m, o, &f);
	if (g <= 0) return g;
	out[0] = 0;
	if (!(f.downlink && f.bs != 0x03)) return 0;                      /* uplinks and squitters never enter the table */
	flight_t *fl = flights_add(t, m, &f, o);
	if (!fl) return ACB_ERR_NOMEM;
	if (o->emptymsg && text_strlen(&f, m->txt) == 0) return 0;
	if (fl->route_sent || !fl->fid[0] || !fl->sa[0] || !fl->da[0]) return 0;       /* routejson, output.c:428-456: once per flight */
	sink_t s = { out, cap, 0, 0 };
	int first = 1;
	putc1(&s, '{');
	json_kv_num(&s, &first, "timestamp", (double)o->tv_sec + ((double)o->tv_usec) / 1e6);
	if (o->station_id && o->station_id[0]) json_kv_str(&s, &first, "station_id", o->station_id, strlen(o->station_id));
	json_kv_str(&s, &first, "flight", fl->fid, 6);
	json_kv_str(&s, &first, "depa", fl->sa, 4);
	json_kv_str(&s, &first, "dsta", fl->da, 4);
	putc1(&s, '}');
	if (s.overflow) return ACB_ERR_ARG;
	fl->route_sent = 1;
	out[s.n] = 0;
	return (int)s.n;
}

/* the monitor screen (printmonitor, output.c:458-484) after this message; `nbch` = channels of the receiver */
int acb_flights_monitor(acb_flights_t *t, const acb_msg_t *m, int nbch, const acb_fmt_opts_t *opt, char *out, size_t cap)
{
	static const acb_fmt_opts_t none = { 0, 0, 0, 0, 0, 0, NULL, NULL };
	if (!t || !m || !out || cap == 0 || nbch < 0 || nbch > 16) return ACB_ERR_ARG;
	const acb_fmt_opts_t *o = opt ? opt : &none;
	acb_fields_t f;
	const int g = flights_gate(m, o, &f);
	if (g <= 0) return g;
	out[0] = 0;
	if (f.downlink && f.bs != 0x03 && !flights_add(t, m, &f, o)) return ACB_ERR_NOMEM;
	if (o->emptymsg && text_strlen(&f, m->txt) == 0) return 0;
	sink_t s = { out, cap, 0, 0 };
	struct tm tmv;
	time_t sec = (time_t)o->tv_sec;
	gmtime_r(&sec, &tmv);
	puts0(&s, "\x1b[H\x1b[2J");
	putf(&s, "             Acarsdec monitor %02d:%02d:%02d.%03ld", tmv.tm_hour, tmv.tm_min, tmv.tm_sec, (long)(o->tv_usec / 1000));
	puts0(&s, "\n Aircraft Flight   Nb Channels     First    DEP   ARR   ETA\n");
	for (const flight_t *fl = t->head; fl; fl = fl->next) {
		putf(&s, " %-8s %-7s %3d ", fl->addr, fl->fid, fl->count);
		int i = 0;
		for (; i < nbch; i++) putc1(&s, (fl->chmask & (1 << i)) ? 'x' : '.');
		for (; i < 16; i++) putc1(&s, ' ');                            /* MAXNBCHANNELS columns */
		sec = (time_t)fl->first_sec;
		gmtime_r(&sec, &tmv);
		putf(&s, " %02d:%02d:%02d.%03ld", tmv.tm_hour, tmv.tm_min, tmv.tm_sec, (long)(fl->first_usec / 1000));
		if (fl->sa[0]) putf(&s, " %4s ", fl->sa); else puts0(&s, "      ");
		if (fl->da[0]) putf(&s, " %4s ", fl->da); else puts0(&s, "      ");
		if (fl->eta[0]) putf(&s, " %4s ", fl->eta); else puts0(&s, "      ");
		putc1(&s, '\n');
	}
	if (s.overflow) return ACB_ERR_ARG;
	out[s.n] = 0;
	return (int)s.n;
}
