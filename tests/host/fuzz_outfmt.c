/* TEST INFRASTRUCTURE: random and malformed blocks through every entry point of csrc/outfmt.c, built with
 * -fsanitize=address,undefined by tests/test_outfmt.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "acars_b200.h"
int main(void)
{
	static char out[16384];
	unsigned long long x = 88172645463325252ull;
	long total = 0;
	acb_flights_t *fl = acb_flights_new(600), *fm = acb_flights_new(30);
	for (int it = 0; it < 80000; it++) {
		acb_msg_t m;
		memset(&m, 0, sizeof m);
		x ^= x << 13; x ^= x >> 7; x ^= x << 17;
		m.len = (int)(x % 260) - 5;               /* also out-of-range lengths */
		m.chn = (int)((x >> 20) % 16);
		m.err = (int)((x >> 30) % 4);
		m.lvl = (float)((x >> 33) % 900) / 10.f - 60.f;
		for (int i = 0; i < ACB_TXTMAX; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; m.txt[i] = (unsigned char)(x & ((it & 3) ? 0x7f : 0xff)); }
		if (it & 1) { m.txt[9] = "Q124813R"[(x >> 5) % 8]; m.txt[10] = "126EB0NZDSTGL"[(x >> 9) % 13]; m.txt[11] = (it & 2) ? '5' : 'B'; m.txt[12] = 2; }
		acb_fmt_opts_t o = { 1700000000 + it * 7, (long)(x % 1000000), 131525000, (int)((x >> 3) % 6), (int)((x >> 11) & 1), (int)((x >> 12) & 1),
		                     (it % 7 == 0) ? "Q1:H1:12" : NULL, (it % 5) ? "STA1" : NULL };
		static const int fmts[] = { 1, 2, 4, 11, 12, 13, 99 };
		for (unsigned k = 0; k < sizeof fmts / sizeof fmts[0]; k++) {
			const int n = acb_format_msg(&m, fmts[k], &o, out, (it % 11 == 0) ? 40 : sizeof out);
			if (n > 0) total += n;
		}
		int n = acb_flights_route_json(fl, &m, &o, out, sizeof out);
		if (n > 0) total += n;
		n = acb_flights_monitor(fm, &m, 8, &o, out, sizeof out);
		if (n > 0) total += n;
	}
	acb_flights_free(fl); acb_flights_free(fm);
	printf("fuzz ok %ld\n", total);
	return 0;
}
