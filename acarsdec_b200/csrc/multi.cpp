/*
 * acb_multi_*: one process, several GPUs (include/acars_b200.h).  Host-side only: it fronts one acb_ctx_t per
 * device and splits either the streams or the channels among them; the devices never talk to each other
 * (channels share nothing but the read-only matched filter, streams share nothing).
 * The reference has no counterpart (one thread serves all channels, rtl.c:344-360); what it fixes for us is the
 * emission order the merged output must keep (rtl.c:357-360: block, then channel, then time).
 */
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <vector>

#include "../../include/acars_b200.h"

struct Part {
	acb_ctx_t *ctx;
	int s0, ns;       /* streams served */
	int c0, nc;       /* channels served */
};

struct acb_multi {
	acb_config_t cfg;
	int mode;
	std::vector<Part> parts;
	std::deque<acb_msg_t> outq;
	std::vector<acb_msg_t> scratch;
};

static void split(int n, int parts, int i, int *lo, int *cnt)
{
	const int base = n / parts, extra = n % parts;
	*lo = i * base + std::min(i, extra);
	*cnt = base + (i < extra ? 1 : 0);
}

extern "C" int acb_multi_create(const acb_config_t *cfg, const int *devices, int ndev, int mode, acb_multi_t **out)
{
	if (!cfg || !devices || !out || ndev < 1) return ACB_ERR_ARG;
	if (mode != ACB_MULTI_SPLIT_STREAMS && mode != ACB_MULTI_SPLIT_CHANNELS) return ACB_ERR_ARG;
	acb_multi *m = new acb_multi();
	m->cfg = *cfg;
	m->mode = mode;
	*out = m;
	const int total = mode == ACB_MULTI_SPLIT_STREAMS ? cfg->nstreams : cfg->nch;
	for (int i = 0; i < ndev; i++) {
		Part p;
		p.ctx = nullptr;
		p.s0 = 0; p.ns = cfg->nstreams; p.c0 = 0; p.nc = cfg->nch;
		if (mode == ACB_MULTI_SPLIT_STREAMS) split(total, ndev, i, &p.s0, &p.ns);
		else split(total, ndev, i, &p.c0, &p.nc);
		if (p.ns == 0 || p.nc == 0) continue;             /* more devices than work: this one stays idle */
		acb_config_t c = *cfg;
		c.device = devices[i];
		c.nstreams = p.ns;
		c.nch = p.nc;
		const int r = acb_create(&c, &p.ctx);
		if (p.ctx) m->parts.push_back(p);
		if (r != ACB_OK) return r;                       /* the caller destroys what exists */
	}
	return m->parts.empty() ? ACB_ERR_ARG : ACB_OK;
}

extern "C" void acb_multi_destroy(acb_multi_t *m)
{
	if (!m) return;
	for (auto &p : m->parts) acb_destroy(p.ctx);
	delete m;
}

extern "C" int acb_multi_parts(acb_multi_t *m) { return m ? (int)m->parts.size() : ACB_ERR_ARG; }
extern "C" acb_ctx_t *acb_multi_part(acb_multi_t *m, int i) { return m && i >= 0 && i < (int)m->parts.size() ? m->parts[i].ctx : nullptr; }

extern "C" int acb_multi_set_plan(acb_multi_t *m, int stream, const unsigned *freqs_hz, int nch, unsigned *fc_out)
{
	if (!m || !freqs_hz || nch != m->cfg.nch || stream < 0 || stream >= m->cfg.nstreams) return ACB_ERR_ARG;
	/* one centre for all channels of the stream (chooseFc, rtl.c:131-168), whoever serves them */
	const unsigned fc = acb_choose_fc(freqs_hz, nch, m->cfg.K);
	if (fc == 0) return ACB_ERR_PLAN;
	if (fc_out) *fc_out = fc;
	for (auto &p : m->parts) {
		if (stream < p.s0 || stream >= p.s0 + p.ns) continue;
		if (int r = acb_set_plan_at(p.ctx, stream - p.s0, freqs_hz + p.c0, p.nc, fc)) return r;
	}
	return ACB_OK;
}

extern "C" int acb_multi_set_wf(acb_multi_t *m, int stream, const float *wf, int nch)
{
	if (!m || !wf || nch != m->cfg.nch || stream < 0 || stream >= m->cfg.nstreams) return ACB_ERR_ARG;
	const int taps = m->cfg.taps ? m->cfg.taps : m->cfg.K;
	for (auto &p : m->parts) {
		if (stream < p.s0 || stream >= p.s0 + p.ns) continue;
		if (int r = acb_set_wf(p.ctx, stream - p.s0, wf + (size_t)p.c0 * taps * 2, p.nc)) return r;
	}
	return ACB_OK;
}

extern "C" int acb_multi_reset(acb_multi_t *m)
{
	if (!m) return ACB_ERR_ARG;
	for (auto &p : m->parts)
		if (int r = acb_reset(p.ctx)) return r;
	m->outq.clear();
	return ACB_OK;
}

extern "C" int acb_multi_submit_host(acb_multi_t *m, const uint8_t *iq, size_t stream_stride, int nblk)
{
	if (!m || !iq) return ACB_ERR_ARG;
	/* asynchronous per device (pinned source): every GPU pulls its part over its own PCIe link at once */
	for (auto &p : m->parts)
		if (int r = acb_submit_host(p.ctx, iq + (size_t)p.s0 * stream_stride, stream_stride, nblk)) return r;
	return ACB_OK;
}

/* pull what every part has queued, translate to global indices, merge in emission order */
static int gather(acb_multi *m)
{
	acb_msg_t buf[64];
	m->scratch.clear();
	for (auto &p : m->parts)
		for (int n; (n = acb_drain(p.ctx, buf, 64)) > 0;)
			for (int i = 0; i < n; i++) {
				buf[i].stream += p.s0;
				buf[i].chn += p.c0;
				m->scratch.push_back(buf[i]);
			}
	std::stable_sort(m->scratch.begin(), m->scratch.end(), [](const acb_msg_t &a, const acb_msg_t &b) {
		if (a.block != b.block) return a.block < b.block;
		if (a.stream != b.stream) return a.stream < b.stream;
		if (a.chn != b.chn) return a.chn < b.chn;
		return a.pos < b.pos;
	});
	for (auto &x : m->scratch) m->outq.push_back(x);
	return (int)m->outq.size();
}

extern "C" int acb_multi_collect(acb_multi_t *m)
{
	if (!m) return ACB_ERR_ARG;
	int err = 0;
	for (auto &p : m->parts) { const int r = acb_collect(p.ctx); if (r < 0 && !err) err = r; }
	const int n = gather(m);
	return err ? err : n;
}

extern "C" int acb_multi_sync(acb_multi_t *m)
{
	if (!m) return ACB_ERR_ARG;
	int err = 0;
	for (auto &p : m->parts) { const int r = acb_sync(p.ctx); if (r < 0 && !err) err = r; }
	const int n = gather(m);
	return err ? err : n;
}

extern "C" int acb_multi_drain(acb_multi_t *m, acb_msg_t *out, int max)
{
	if (!m || (!out && max > 0)) return ACB_ERR_ARG;
	int n = 0;
	while (n < max && !m->outq.empty()) { out[n++] = m->outq.front(); m->outq.pop_front(); }
	return n;
}

static Part *find(acb_multi *m, int stream, int chn)
{
	for (auto &p : m->parts)
		if (stream >= p.s0 && stream < p.s0 + p.ns && chn >= p.c0 && chn < p.c0 + p.nc) return &p;
	return nullptr;
}

extern "C" int acb_multi_get_state(acb_multi_t *m, int stream, int chn, acb_chan_state_t *out)
{
	Part *p = m ? find(m, stream, chn) : nullptr;
	return p ? acb_get_state(p->ctx, stream - p->s0, chn - p->c0, out) : ACB_ERR_ARG;
}

extern "C" int acb_multi_set_state(acb_multi_t *m, int stream, int chn, const acb_chan_state_t *in)
{
	Part *p = m ? find(m, stream, chn) : nullptr;
	return p ? acb_set_state(p->ctx, stream - p->s0, chn - p->c0, in) : ACB_ERR_ARG;
}
