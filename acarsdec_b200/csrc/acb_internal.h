/* Internal device/host shared layouts of libacars_b200 (not part of the C ABI). */
#ifndef ACB_INTERNAL_H
#define ACB_INTERNAL_H

#include <stdint.h>

namespace acb {

constexpr int FLEN = 11;        /* msk.c:25  INTRATE/1200 + 1 */
constexpr int MFLTOVER = 12;    /* msk.c:26 */
constexpr int FLENO = 133;      /* msk.c:27 */
constexpr int OUTBLK = 1024;    /* rtl.c:49 */
constexpr int TXTCAP = 256;     /* >= acarsdec.h:55's 250, padded */

/* Persistent per-(stream,channel) demodulator + framing state in HBM: what channel_t carries
 * between demodMSK calls (acarsdec.h:76-89; msk.c:71-72,134-135), plus the frame being
 * assembled (msgblk_t, acarsdec.h:48-57).  8-byte aligned, 448 bytes. */
struct ChainState {
	double phi, df, lvlsum;           /* MskPhi, MskDf, MskLvlSum */
	float clk;                        /* MskClk (float state!) */
	int bitcount;                     /* MskBitCount */
	unsigned S, idx;                  /* MskS, idx */
	int nbits, state;                 /* nbits, Acarsstate */
	unsigned outbits;
	int blk_len;
	int blk_err;
	int pad0;
	unsigned long long pos;           /* envelope samples consumed so far */
	unsigned long long soh_pos;       /* pos at the SOH of the frame in progress */
	float inb_re[FLEN], inb_im[FLEN]; /* msk.c:40 ring */
	unsigned char crc[2];
	unsigned char pad1[6];
	unsigned char txt[TXTCAP];
};

/* A completed (pre-FEC) frame, device -> host. */
struct RawFrame {
	int stream, chn, len, err;
	double lvlsum;
	int bitcount;
	int pad0;                         /* block-FEC status: 0 = not processed, 1 = deliver, 2 = drop */
	unsigned long long pos, soh_pos;
	unsigned char crc[2];
	unsigned char pad1[6];
	unsigned char txt[TXTCAP];
};

struct RingCtl {
	unsigned count;                   /* frames appended (may exceed capacity: overflow) */
	unsigned short_frames;            /* frames of fewer than 13 bytes, dropped before taking a slot (acars.c:124-129) */
	unsigned pad[2];
};

/* channelizer tile geometry */
constexpr int CH_TILE = 128;          /* threads per channelizer CTA (2 output rows each) */
constexpr int CH_GROUP = 8;           /* channels accumulated per pass */

} // namespace acb

/* kernels.cu entry points (host-callable launchers) */
struct CUstream_st;
namespace acb {
enum InputKind { IN_KIND_U8IQ = 0, IN_KIND_F32REAL = 1, IN_KIND_CS16IQ = 2 };
int launch_channelize(int mode, const void *in, size_t stream_stride_bytes, const float *wf, float *dm,
                      int K, int taps, int nch, int nstreams, int nblk, size_t nsamp, CUstream_st *stream);
int launch_channelize_generic(int mode, const void *in, size_t stream_stride, const void *wf, float *dm,
                              int K, int taps, int nch, int nstreams, size_t row0, size_t nrows, size_t nsamp, CUstream_st *stream);
int launch_demod(ChainState *st, const float *dm, int nsamp, int nch, int nstreams,
                 RawFrame *ring, RingCtl *ctl, unsigned cap, int lanes_per_channel, CUstream_st *stream);
int demod_pick_lanes(long long nchains, int sm_count);
bool channelize_dft_supports(int K);
int launch_channelize_dft(const void *in, size_t stream_stride, const float *tw, const unsigned *meta, float *dm,
                          int K, int nch, int nstreams, int nblk, size_t nsamp, bool fold8, CUstream_st *stream);
int launch_channelize_dft_cs16(const void *in, size_t stream_stride, const float *tw, const unsigned *meta, float *dm,
                               int K, int nch, int nstreams, size_t nsamp, CUstream_st *stream);
bool channelize_rdft_supports(int K);
int launch_channelize_rdft(const void *in, size_t stream_stride, const float *tw, const unsigned *meta, float *dm,
                           int K, int nch, int nstreams, size_t nsamp, CUstream_st *stream);
int launch_block_fec(RawFrame *ring, const RingCtl *ctl, unsigned cap, CUstream_st *stream);
int launch_interleave_cs16(const int16_t *xi, const int16_t *xq, size_t plane_stride, uint32_t *out, size_t out_stride,
                           size_t nsamples, int nstreams, CUstream_st *stream);
int upload_matched_filter(const float *h, CUstream_st *stream);
int upload_sincos_table(const double *cos_hi_lo, const double *sin_hi_lo, CUstream_st *stream);
size_t channelize_smem_bytes(int mode);
} // namespace acb

#endif
