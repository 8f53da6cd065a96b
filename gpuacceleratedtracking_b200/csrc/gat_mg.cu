// gat_mg.cu -- gat_mg_*: one host process (the Julia tracking loop) driving the GPUs of one box through ONE call.
//
// `gat_mg_create`, `gat_mg_correlate` "(same args; does broadcast + shard + gather)".  The reference is
// single-device (src/benchmarks.jl:24) and only hints at the channel axis (`sat_idx = blockIdx().z`, src/algorithms.jl:656;
// paper/paper.tex:114), so this layer is specified by north_star: satellite channels are partitioned over the GPUs, every
// signal block reaches every GPU, only the small accumulators come back to the host.
//   * the block exchange is the signal ring (gat_ring.cu): gat_mg_upload_signal sends each device ITS sample range of the
//     block through that device's own PCIe link; the correlate kernels gather the other ranges over NVLink tile by tile;
//   * channels are sorted by system id (bands stay together) and cut into contiguous, balanced shards;
//   * every device's launch is queued before any result is awaited, so the GPUs run concurrently from one host thread;
//     the accumulators land in the caller's arrays in the ORIGINAL channel order.
//   * DEFAULT since round 2's second session: SAMPLE sharding (gat_mg_set_sharding).  The accumulators are sums over samples,
//     so every device correlates ALL channels over the sample range it already holds (its share of the ring) and the host adds
//     the devices' partial sums in device order: no signal crosses NVLink at all, 8 K L M bytes of partial sums per period
//     come back per device.  Measured on 8 x B200 (bench.py c5 leg): C5 weak 35 us -> see DESIGN.md section 6.  The phases are
//     taken with the whole-block call's integer arithmetic (gat_set_sample_origin's mechanism), so the chip indices are
//     bit-exact; the sums differ from a single-device call in FP32 summation order only.
#include <new>
#include <numeric>

#include "gat_ctx.h"

using namespace gat;

struct gat_mg {
    std::vector<gat_ctx *> ctx;
    std::string err;
    int n_slots = 0, n_samples = 0, n_ants = 0;
    bool dirty = false;                    // uploads queued since the last publish
    int generation = 0;
    std::vector<int> slot_release;         // release count after the last correlate that read the slot
    int releases = 0;
    std::vector<float *> d_out;            // per device: re | im accumulators of its shard
    std::vector<size_t> d_out_cap;
    std::vector<float> h_tmp;
    int sharding = 0;                      // 0 = samples (default), 1 = satellites (signal exchange over NVLink)
};
constexpr int kMgLocalSlotBase = 60000;    // plain slots bound to each device's own share of the ring slots

namespace {

int mg_fail(gat_mg *mg, int status, const std::string &msg)
{
    if (mg) mg->err = msg;
    return status;
}

int mg_ctx_fail(gat_mg *mg, int i, int status)
{
    return mg_fail(mg, status, "device " + std::to_string(i) + ": " + gat_last_error(mg->ctx[i]));
}

#define MG_CUDA(mg, i, call)                                                                          \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return mg_fail(mg, GAT_ERR_CUDA, "device " + std::to_string(i) + ": " #call ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

extern "C" {

int gat_mg_destroy(gat_mg *mg)
{
    if (!mg) return GAT_ERR_INVALID;
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        if (!mg->ctx[i]) continue;
        cudaSetDevice(mg->ctx[i]->device);
        cudaStreamSynchronize(mg->ctx[i]->stream);
        cudaStreamSynchronize(mg->ctx[i]->copy_stream);
    }
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        if (!mg->ctx[i]) continue;
        cudaSetDevice(mg->ctx[i]->device);
        if (i < mg->d_out.size() && mg->d_out[i]) cudaFree(mg->d_out[i]);
        gat_destroy(mg->ctx[i]);
    }
    delete mg;
    return GAT_OK;
}

int gat_mg_create(gat_mg **out, int n_dev, const int *devices)
{
    if (!out || n_dev < 1 || n_dev > kMaxPeers || !devices) return GAT_ERR_INVALID;
    *out = nullptr;
    gat_mg *mg = new (std::nothrow) gat_mg();
    if (!mg) return GAT_ERR_INVALID;
    mg->ctx.assign(n_dev, nullptr);
    mg->d_out.assign(n_dev, nullptr);
    mg->d_out_cap.assign(n_dev, 0);
    for (int i = 0; i < n_dev; ++i) {
        int rc = gat_create(&mg->ctx[i], devices[i]);
        if (rc) {
            gat_mg_destroy(mg);
            return rc;
        }
    }
    *out = mg;
    return GAT_OK;
}

const char *gat_mg_last_error(gat_mg *mg) { return mg ? mg->err.c_str() : "null multi-GPU context"; }
int gat_mg_device_count(gat_mg *mg) { return mg ? static_cast<int>(mg->ctx.size()) : GAT_ERR_INVALID; }
gat_ctx *gat_mg_ctx(gat_mg *mg, int i) { return (mg && i >= 0 && i < static_cast<int>(mg->ctx.size())) ? mg->ctx[i] : nullptr; }

int gat_mg_set_codes(gat_mg *mg, int system_id, const int8_t *chips, int code_len, int n_prn)
{
    if (!mg) return GAT_ERR_INVALID;
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        int rc = gat_set_codes(mg->ctx[i], system_id, chips, code_len, n_prn);
        if (rc) return mg_ctx_fail(mg, static_cast<int>(i), rc);
    }
    return GAT_OK;
}

int gat_mg_configure(gat_mg *mg, int n_slots, int n_samples, int n_ants)
{
    if (!mg) return GAT_ERR_INVALID;
    const int world = static_cast<int>(mg->ctx.size());
    unsigned char handle[GAT_IPC_HANDLE_BYTES];
    for (int i = 0; i < world; ++i) {
        int rc = gat_ring_create(mg->ctx[i], world, i, n_slots, n_samples, n_ants, handle);
        if (rc) return mg_ctx_fail(mg, i, rc);
    }
    for (int i = 0; i < world; ++i) {
        int rc = gat_ring_connect_local(mg->ctx[i], mg->ctx.data());
        if (rc) return mg_ctx_fail(mg, i, rc);
    }
    mg->n_slots = n_slots;
    mg->n_samples = n_samples;
    mg->n_ants = n_ants;
    mg->dirty = false;
    mg->generation = 0;
    mg->releases = 0;
    mg->slot_release.assign(n_slots, 0);
    // sample sharding reads every device's OWN share of a ring slot through a plain slot bound to that memory
    for (int i = 0; i < world; ++i) {
        gat_ctx *c = mg->ctx[i];
        for (int s = 0; s < n_slots; ++s) {
            const SignalSlot *rs = find_slot(c, s);
            if (!rs || static_cast<int>(rs->parts.size()) <= i || rs->parts[i].len < 1) continue;
            const SlotPart &lp = rs->parts[i];
            int rc = gat_bind_signal(c, kMgLocalSlotBase + s, lp.re, lp.im, lp.len, n_ants, static_cast<int>(lp.ld));
            if (rc) return mg_ctx_fail(mg, i, rc);
        }
    }
    return GAT_OK;
}

int gat_mg_set_sharding(gat_mg *mg, int mode)
{
    if (!mg || mode < 0 || mode > 1) return GAT_ERR_INVALID;
    mg->sharding = mode;
    return GAT_OK;
}

int gat_mg_upload_signal(gat_mg *mg, int slot, const float *h_re, const float *h_im, int ld)
{
    NvtxRange nvtx_call("gat_mg_upload_signal");
    if (!mg || !mg->n_slots) return mg_fail(mg, GAT_ERR_INVALID, "gat_mg_configure first");
    if (slot < 0 || slot >= mg->n_slots || !h_re || !h_im || ld < mg->n_samples) return mg_fail(mg, GAT_ERR_INVALID, "bad upload arguments");
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        // the slot may still be read by correlate calls queued earlier (on ANY device): hold this device's ingest stream
        // until every device has released them
        int rc = gat_ring_acquire(mg->ctx[i], mg->slot_release[slot]);
        if (!rc) rc = gat_ring_upload(mg->ctx[i], slot, h_re, h_im, ld, 0);
        if (rc) return mg_ctx_fail(mg, static_cast<int>(i), rc);
    }
    mg->dirty = true;
    return GAT_OK;
}

int gat_mg_correlate(gat_mg *mg, int n_periods, const int32_t *slots, int n_sats, const gat_channel *channels, double fs_hz,
                     const int32_t *sample_shifts, int n_taps, int start_sample, int n_samples, float *h_out_re, float *h_out_im,
                     unsigned flags)
{
    NvtxRange nvtx_call("gat_mg_correlate");
    if (!mg || !mg->n_slots) return mg_fail(mg, GAT_ERR_INVALID, "gat_mg_configure first");
    if (!slots || !channels || !sample_shifts || !h_out_re || !h_out_im || n_periods < 1 || n_sats < 1 || n_taps < 1 || n_taps > GAT_MAX_TAPS)
        return mg_fail(mg, GAT_ERR_INVALID, "bad correlate arguments");
    if (flags & (GAT_ACCUMULATE | GAT_GATHER | GAT_TENSOR_TF32))
        return mg_fail(mg, GAT_ERR_INVALID, "gat_mg_correlate returns host results from ring slots: no ACCUMULATE / GATHER / TENSOR");
    for (int p = 0; p < n_periods; ++p)
        if (slots[p] < 0 || slots[p] >= mg->n_slots) return mg_fail(mg, GAT_ERR_INVALID, "slot outside the configured ring");
    const int world = static_cast<int>(mg->ctx.size());
    const int M = mg->n_ants;
    if (mg->sharding == 0) {
        // ---- sample sharding: every device, all channels, its own sample range; the host adds the partial sums ----
        if (start_sample < 0 || n_samples < 1 || static_cast<int64_t>(start_sample) + n_samples > mg->n_samples)
            return mg_fail(mg, GAT_ERR_INVALID, "sample range outside the configured blocks");
        if (mg->dirty) {
            for (int i = 0; i < world; ++i) {
                int g = gat_ring_publish(mg->ctx[i]);
                if (g < 0) return mg_ctx_fail(mg, i, g);
                mg->generation = g;
            }
            mg->dirty = false;
        }
        const size_t elems = static_cast<size_t>(n_periods) * n_sats * n_taps * M;
        std::vector<int32_t> loc(n_periods);
        for (int p = 0; p < n_periods; ++p) loc[p] = kMgLocalSlotBase + slots[p];
        std::vector<char> active(world, 0);
        for (int i = 0; i < world; ++i) {
            gat_ctx *c = mg->ctx[i];
            const Ring &rg = c->ring;
            const int ps = rg.part_start[i], pl = rg.part_len[i];
            const int a0 = std::max(start_sample, ps), a1 = std::min(start_sample + n_samples, ps + pl);
            MG_CUDA(mg, i, cudaSetDevice(c->device));
            int rc = gat_ring_wait(c, mg->generation);          // this device's share of the block has landed (ingest stream)
            if (rc) return mg_ctx_fail(mg, i, rc);
            if (a1 <= a0) continue;                             // the range does not touch this device's share
            if (2 * elems > mg->d_out_cap[i]) {
                MG_CUDA(mg, i, cudaStreamSynchronize(c->stream));
                if (mg->d_out[i]) MG_CUDA(mg, i, cudaFree(mg->d_out[i]));
                mg->d_out[i] = nullptr;
                mg->d_out_cap[i] = 0;
                MG_CUDA(mg, i, cudaMalloc(reinterpret_cast<void **>(&mg->d_out[i]), 4 * elems * sizeof(float)));
                mg->d_out_cap[i] = 4 * elems;
            }
            // the caller's phases refer to start_sample: slot sample s of this share is sample ps + s - start_sample of that frame
            c->sample_origin_on = true;
            c->sample_origin = ps - start_sample;
            rc = gat_correlate_batch(c, n_periods, loc.data(), n_sats, channels, fs_hz, sample_shifts, n_taps, a0 - ps, a1 - a0, mg->d_out[i],
                                     mg->d_out[i] + elems, 1, flags);
            c->sample_origin_on = false;
            c->sample_origin = 0;
            if (rc) return mg_ctx_fail(mg, i, rc);
            active[i] = 1;
        }
        for (int i = 0; i < world; ++i) {
            int r = gat_ring_release(mg->ctx[i]);
            if (r < 0) return mg_ctx_fail(mg, i, r);
            mg->releases = r;
        }
        for (int p = 0; p < n_periods; ++p) mg->slot_release[slots[p]] = mg->releases;
        std::fill(h_out_re, h_out_re + elems, 0.f);
        std::fill(h_out_im, h_out_im + elems, 0.f);
        mg->h_tmp.resize(2 * elems);
        for (int i = 0; i < world; ++i) {                       // fixed device order: bit-reproducible sums
            if (!active[i]) continue;
            gat_ctx *c = mg->ctx[i];
            MG_CUDA(mg, i, cudaSetDevice(c->device));
            MG_CUDA(mg, i, cudaMemcpyAsync(mg->h_tmp.data(), mg->d_out[i], 2 * elems * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            MG_CUDA(mg, i, cudaStreamSynchronize(c->stream));
            for (size_t e = 0; e < elems; ++e) {
                h_out_re[e] += mg->h_tmp[e];
                h_out_im[e] += mg->h_tmp[elems + e];
            }
        }
        return GAT_OK;
    }
    // ---- satellite sharding ----
    // shard the channel axis: stable order by system id (a device then tends to need one band's tables), contiguous and balanced
    std::vector<int> order(n_sats);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return channels[a].system_id < channels[b].system_id; });
    if (mg->dirty) {
        for (int i = 0; i < world; ++i) {
            int g = gat_ring_publish(mg->ctx[i]);
            if (g < 0) return mg_ctx_fail(mg, i, g);
            mg->generation = g;
        }
        mg->dirty = false;
    }
    std::vector<int> lo(world + 1, 0);
    for (int i = 0; i < world; ++i) lo[i + 1] = lo[i] + n_sats / world + (i < n_sats % world ? 1 : 0);
    std::vector<std::vector<gat_channel>> shard(world);
    const size_t per_ch = static_cast<size_t>(n_taps) * M;
    // queue every device's work first ...
    for (int i = 0; i < world; ++i) {
        const int K = lo[i + 1] - lo[i];
        if (K == 0) continue;
        gat_ctx *c = mg->ctx[i];
        shard[i].resize(static_cast<size_t>(n_periods) * K);
        for (int p = 0; p < n_periods; ++p)
            for (int k = 0; k < K; ++k) shard[i][static_cast<size_t>(p) * K + k] = channels[static_cast<size_t>(p) * n_sats + order[lo[i] + k]];
        const size_t elems = static_cast<size_t>(n_periods) * K * per_ch;
        MG_CUDA(mg, i, cudaSetDevice(c->device));
        if (2 * elems > mg->d_out_cap[i]) {
            MG_CUDA(mg, i, cudaStreamSynchronize(c->stream));
            if (mg->d_out[i]) MG_CUDA(mg, i, cudaFree(mg->d_out[i]));
            mg->d_out[i] = nullptr;
            mg->d_out_cap[i] = 0;
            MG_CUDA(mg, i, cudaMalloc(reinterpret_cast<void **>(&mg->d_out[i]), 4 * elems * sizeof(float)));
            mg->d_out_cap[i] = 4 * elems;
        }
        int rc = gat_ring_wait(c, mg->generation);
        if (!rc)
            rc = gat_correlate_batch(c, n_periods, slots, K, shard[i].data(), fs_hz, sample_shifts, n_taps, start_sample, n_samples,
                                     mg->d_out[i], mg->d_out[i] + elems, 1, flags);
        if (rc) return mg_ctx_fail(mg, i, rc);
    }
    // (every device releases, also one without channels: the flag protocol counts all ranks)
    for (int i = 0; i < world; ++i) {
        int r = gat_ring_release(mg->ctx[i]);
        if (r < 0) return mg_ctx_fail(mg, i, r);
        mg->releases = r;
    }
    for (int p = 0; p < n_periods; ++p) mg->slot_release[slots[p]] = mg->releases;
    // ... then collect: device i's [M x L x K_i x P] block goes to columns order[lo_i ..] of the caller's [M x L x K x P]
    for (int i = 0; i < world; ++i) {
        const int K = lo[i + 1] - lo[i];
        if (K == 0) continue;
        gat_ctx *c = mg->ctx[i];
        const size_t elems = static_cast<size_t>(n_periods) * K * per_ch;
        mg->h_tmp.resize(2 * elems);
        MG_CUDA(mg, i, cudaSetDevice(c->device));
        MG_CUDA(mg, i, cudaMemcpyAsync(mg->h_tmp.data(), mg->d_out[i], 2 * elems * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        MG_CUDA(mg, i, cudaStreamSynchronize(c->stream));
        for (int p = 0; p < n_periods; ++p)
            for (int k = 0; k < K; ++k) {
                const size_t src = (static_cast<size_t>(p) * K + k) * per_ch;
                const size_t dst = (static_cast<size_t>(p) * n_sats + order[lo[i] + k]) * per_ch;
                std::memcpy(h_out_re + dst, mg->h_tmp.data() + src, per_ch * sizeof(float));
                std::memcpy(h_out_im + dst, mg->h_tmp.data() + elems + src, per_ch * sizeof(float));
            }
    }
    return GAT_OK;
}

int gat_mg_sync(gat_mg *mg)
{
    if (!mg) return GAT_ERR_INVALID;
    for (size_t i = 0; i < mg->ctx.size(); ++i) {
        int rc = gat_sync(mg->ctx[i]);
        if (rc) return mg_ctx_fail(mg, static_cast<int>(i), rc);
    }
    return GAT_OK;
}

}  // extern "C"
