/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Single-lane host build of the demodulator loop the CUDA kernel k_demod2 runs
 * (acarsdec_b200/csrc/demod_core.h), so that the CPU test suite can hold the loop's control flow and
 * arithmetic (fast path / general path selection, the one-division decision, the phase-index guard,
 * the doubled ring) to the oracle bit for bit without a GPU.  The lane plumbing of the kernel (who
 * evaluates which mixer sample) is covered by the -m gpu tests.  Built by tests/test_demod_core.py with
 * g++ -O2 -ffp-contract=off; links nothing from the product but the table builders of libacars_b200.so.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/acars_b200.h"
#include "../../acarsdec_b200/csrc/demod_core.h"

using namespace acb;

struct EmulFrame {
	int len, err, bitcount, pad;
	double lvlsum;
	uint64_t pos, soh_pos;
	unsigned char crc[2];
	unsigned char txt[ACB_TXTMAX];
};

struct HostEnv {
	static bool any(bool x) { return x; }
	static bool all(bool x) { return x; }
	static void sync() {}
	static int max(int x) { return x; }
};

struct HostFrameAcc {
	DemodRegs &r;
	acb_chan_state_t *st;
	EmulFrame *out;
	int max, n;
	int &state() { return r.state; }
	int &nbits() { return r.nbits; }
	int &bitcount() { return r.bitcount; }
	int &blk_len() { return r.blk_len; }
	int &blk_err() { return r.blk_err; }
	unsigned &msk_s() { return r.S; }
	double &msk_df() { return r.df; }
	double &lvlsum() { return r.lvlsum; }
	void txt_put(int i, unsigned char c) { st->blk_txt[i] = c; }
	unsigned char txt_get(int i) { return st->blk_txt[i]; }
	void crc_put(int i, unsigned char c) { st->blk_crc[i] = c; }
	bool frame_begin() { r.soh_pos = r.pos0 + (unsigned long long)(long long)r.fire_n; return true; }
	void frame_emit()
	{
		if (n < max) {
			EmulFrame &f = out[n];
			memset(&f, 0, sizeof(f));
			f.len = r.blk_len; f.err = r.blk_err; f.bitcount = r.bitcount; f.lvlsum = r.lvlsum;
			f.pos = r.pos0 + (unsigned long long)(long long)r.fire_n; f.soh_pos = r.soh_pos;
			f.crc[0] = st->blk_crc[0]; f.crc[1] = st->blk_crc[1];
			memcpy(f.txt, st->blk_txt, ACB_TXTMAX);
		}
		n++;
	}
};

extern "C" void acb_build_sincos_table(double *cos_hi_lo, double *sin_hi_lo);

/* one demodMSK call (msk.c:67) over dm[0], dm[stride], ...; returns the frames completed (pre-FEC) */
extern "C" int demod_emul(acb_chan_state_t *st, const float *dm, int nsamp, int stride, int f2f, EmulFrame *frames, int maxframes)
{
	static DemodShared<1> sm;
	static bool init = false;
	if (!init) {
		float h[FLENO];
		acb_build_h(h);
		for (int o = 0; o <= MFLTOVER; o++) {
			float t[12];
			for (int j = 0; j < 12; j++) t[j] = j < FLEN ? h[o + MFLTOVER * j] : 0.f;
			for (int q = 0; q < 3; q++) { sm.h2[o][q].x = t[4 * q]; sm.h2[o][q].y = t[4 * q + 1]; sm.h2[o][q].z = t[4 * q + 2]; sm.h2[o][q].w = t[4 * q + 3]; }
		}
		double tc[128], ts[128];
		acb_build_sincos_table(tc, ts);
		for (int k = 0; k < 64; k++) { sm.tcos[k].x = tc[2 * k]; sm.tcos[k].y = tc[2 * k + 1]; sm.tsin[k].x = ts[2 * k]; sm.tsin[k].y = ts[2 * k + 1]; }
		init = true;
	}
	DemodRegs r;
	r.phi = st->MskPhi; r.df = st->MskDf; r.lvlsum = st->MskLvlSum; r.clk = st->MskClk; r.bitcount = st->MskBitCount;
	r.S = st->MskS; r.idx = st->idx % FLEN; r.nbits = st->nbits; r.state = st->Acarsstate; r.outbits = st->outbits;
	r.blk_len = st->blk_len; r.blk_err = st->blk_err; r.pos = st->pos; r.soh_pos = st->soh_pos;
	for (int k = 0; k < FLEN; k++) {
		DcF2 v;
		v.x = st->inb_re[k]; v.y = st->inb_im[k];
		sm.ring[k][0] = v;
		sm.ring[k + FLEN][0] = v;
	}
	HostFrameAcc acc{ r, st, frames, maxframes, 0 };
	if (f2f) demod_run<1, true, true, HostEnv>(r, sm, dm, stride, nsamp, 0, 0, acc);
	else demod_run<1, false, true, HostEnv>(r, sm, dm, stride, nsamp, 0, 0, acc);
	st->MskPhi = r.phi; st->MskDf = r.df; st->MskLvlSum = r.lvlsum; st->MskClk = r.clk; st->MskBitCount = r.bitcount;
	st->MskS = r.S; st->idx = r.idx; st->nbits = r.nbits; st->Acarsstate = r.state; st->outbits = r.outbits;
	st->blk_len = r.blk_len; st->blk_err = r.blk_err; st->pos = r.pos; st->soh_pos = r.soh_pos;
	for (int k = 0; k < FLEN; k++) { st->inb_re[k] = sm.ring[k][0].x; st->inb_im[k] = sm.ring[k][0].y; }
	return acc.n;
}
