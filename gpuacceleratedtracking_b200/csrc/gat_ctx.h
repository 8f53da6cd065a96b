// gat_ctx.h -- the context object behind include/gat.h's opaque gat_ctx and the small host helpers every
// translation unit of the C-ABI layer shares (gat_api.cu, gat_ring.cu, gat_mg.cu).  Internal.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/gat.h"
#include "gat_internal.h"

namespace gat {

// NVTX range around every host-side stage (the reference wraps each launch in `NVTX.@range`,
// src/algorithms.jl:953, :973, :1015 ... :1526, and profiles with scripts/nsys.jl:100): free when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

// One owner's share of a sharded slot (gat_ring_*): samples [start, start + len) of every antenna, in that owner's
// HBM -- local memory, or a peer mapping (CUDA IPC / direct peer access) the kernel TMA-loads from over NVLink.
struct SlotPart {
    float *re = nullptr, *im = nullptr;
    int64_t ld = 0;
    int start = 0, len = 0;
    PeriodDev maps{};
};

struct SignalSlot {
    float *re = nullptr, *im = nullptr;
    int64_t ld = 0;
    int n_samples = 0, n_ants = 0;
    bool owned = false;
    size_t cap_floats = 0;  // per plane, when owned
    PeriodDev maps{};       // TMA descriptors of the two planes
    bool maps_valid = false;
    bool planes_valid = false;   // re / im hold the block (false while only the raw integer copy exists)
    // raw interleaved complex int16 copy of the block (gat_upload_signal_sc16): [n_ants][raw_ld] words of I | Q << 16
    int16_t *raw = nullptr;
    size_t raw_cap = 0;          // complex samples
    int64_t raw_ld = 0;
    float raw_scale = 1.f;
    bool raw_valid = false;
    PeriodDev raw_map{};         // .re = 2-D descriptor over the 32-bit I/Q words
    void *peer_base = nullptr;   // gat_slot_import: another process's planes mapped through CUDA IPC (closed on release)
    TcPeriod tc_map{};           // 4-D descriptor of both planes for the tensor-core path (encoded on first use)
    int tc_state = 0;            // 0 = not tried, 1 = valid, -1 = the layout cannot be expressed (im <= re, ...)
    std::vector<SlotPart> parts; // non-empty: a sharded slot of the signal ring; re / im / raw are unused
    int part_tiles = 0;          // tiles (of kTileCap samples) per part
};

// The signal ring of gat_ring_*: `n_slots` blocks whose samples are spread over the `world` ranks' HBM.  Rank r's
// allocation = [256 B of flags: pub[8] | rel[8]] [slot 0: re plane, im plane] [slot 1] ..., plane = n_ants x part_ld[r].
constexpr int kRingEvents = 16;
struct Ring {
    int world = 0, rank = 0, n_slots = 0, n_samples = 0, n_ants = 0;
    int part_tiles = 0, n_parts = 0;            // parts that hold samples (the leading ranks when the block is short)
    int part_start[kMaxPeers] = {}, part_len[kMaxPeers] = {};
    int64_t part_ld[kMaxPeers] = {};
    unsigned char *base[kMaxPeers] = {};        // rank r's allocation as this process addresses it
    void *opened[kMaxPeers] = {};               // mappings to close (cudaIpcOpenMemHandle)
    unsigned char *local = nullptr;
    unsigned int pub_seq = 0, rel_seq = 0;      // generations published / released by this rank so far
    bool connected = false;
    // mirror view (gat_ring_enable_mirror): local copies of the peers' shares, refreshed by copy-engine prefetches
    unsigned char *mirror[kMaxPeers] = {};
    cudaEvent_t rel_ev[kRingEvents] = {}, pf_ev[kRingEvents] = {};
    int pf_tickets = 0;
    bool mirrored = false;
};
constexpr size_t kRingFlagBytes = 256;
constexpr int kIngestDepth = 3, kIngestChunk = 16, kIngestSlotBase = 65000;   // gat_ingest_correlate staging

struct CodeTable {
    int8_t *d_chips = nullptr;   // [n_prn][col_stride], columns zero-padded to kCodeColAlign bytes
    int code_len = 0, n_prn = 0, col_stride = 0;
    double code_freq_hz = 0.0;   // nominal chip rate (gat_gen_signal); 0 = unknown (caller table without gat_set_code_frequency)
};

struct Staging {
    unsigned char *h = nullptr;  // pinned
    unsigned char *d = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;      // H2D copy finished (param stream)
    cudaEvent_t consumed = nullptr;  // the kernel reading `d` was queued behind this (main stream)
    bool pending = false;
};

constexpr int kStagingRing = 8;

// gat_resident_*: one session = one shape, a set of slots, a kernel that stays on the device (gat_resident.cu)
struct Resident {
    bool active = false;
    cudaStream_t stream = nullptr;      // the resident kernel's own (non-blocking) stream
    LaunchPlan plan{};
    CorrArgs args{};
    ResCtl ctl{};
    size_t smem = 0;
    uint32_t *h_cmd = nullptr;          // pinned, mapped: kResMaxCells cells of 4 words
    uint32_t *h_flag = nullptr;         // pinned, mapped: debug stamps live behind it
    uint4 *d_relay = nullptr;
    unsigned long long *h_res = nullptr;   // pinned, mapped: [re block | im block] of {value bits, sequence number} words
    PeriodDev *d_maps = nullptr;
    uint32_t seq = 0;                   // last command issued
    bool launched = false;              // a kernel was launched and has not been seen finished
    int n_slots = 0, n_sats = 0, n_taps = 0, n_ants = 0, rep_len = 0;
    double fs_hz = 0.0;
    size_t out_elems = 0;
    uint64_t relaunches = 0;
    // GAT_RESIDENT_DEBUG: where a call's time goes (host clock around the call, device stamps inside it)
    bool debug = false;
    double dbg_host_ns = 0, dbg_seen_to_body_ns = 0, dbg_body_ns = 0;
    uint64_t dbg_calls = 0;
};

}  // namespace gat

struct gat_ctx {
    int device = 0;
    int n_sm = 0;
    int max_ctas = 0;   // gat_set_max_ctas: 0 = one CTA on every SM
    cudaStream_t stream = nullptr;      // the stream work is queued on
    cudaStream_t own_stream = nullptr;  // created by gat_create
    cudaStream_t param_stream = nullptr;  // parameter-block uploads, overlapping the previous kernel
    cudaStream_t copy_stream = nullptr;   // signal ingest (gat_ring_upload*, publish, acquire), overlapping the kernels
    gat::Ring ring;
    std::string err;
    gat::CodeTable codes[GAT_MAX_SYSTEMS];
    std::map<int, gat::SignalSlot> slots;   // by slot id (sparse: ids up to 65535 cost nothing until used)
    gat::Staging stg[gat::kStagingRing];
    int stg_next = 0;
    float *d_partials = nullptr;
    size_t partials_cap = 0;
    unsigned int *d_barrier = nullptr;   // grid-barrier arrival counter (monotonic)
    unsigned int barrier_count = 0;      // host mirror: value after all launches queued so far
    float *d_out = nullptr;
    size_t d_out_cap = 0;
    float *h_out = nullptr;  // pinned
    size_t h_out_cap = 0;
    int32_t *d_dbg = nullptr;
    size_t d_dbg_cap = 0;
    int dump_tiles = 0, dump_stride = 0, dump_tile_len = 0, dump_aligned_start = 0;   // geometry of the last replica-index dump
    unsigned char *d_raw = nullptr;      // raw integer samples awaiting expansion
    size_t d_raw_cap = 0;
    // fused multi-GPU gather
    unsigned char *g_local = nullptr;          // this rank's allocation: re | im | flags
    void *g_opened[gat::kMaxPeers] = {};            // peer base pointers from cudaIpcOpenMemHandle
    float *g_re[gat::kMaxPeers] = {}, *g_im[gat::kMaxPeers] = {};
    unsigned int *g_flag[gat::kMaxPeers] = {};
    uint64_t g_elems = 0, g_off = 0;
    int g_world = 0, g_rank = 0;
    bool g_connected = false;
    unsigned int g_seq = 0;
    unsigned int *d_done = nullptr;
    unsigned long long *d_timeline = nullptr;
    size_t timeline_cap = 0;
    bool timeline_on = false;
    int timeline_ctas = 0;
    gat_launch_info info{};
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    uint64_t launches = 0;
    // gat_ingest_correlate: host blocks -> staging slots (copy stream) -> kernel, chunk by chunk
    cudaEvent_t ing_ready[gat::kIngestDepth] = {}, ing_free[gat::kIngestDepth] = {};
    float *d_ing_out = nullptr;
    size_t ing_out_cap = 0;
    float *d_ing_stage[gat::kIngestDepth] = {};      // per ring buffer: re planes of a chunk, then its im planes (contiguous)
    size_t ing_stage_cap = 0;                        // floats per ring buffer
    int ing_n = 0, ing_m = 0;                        // shape the staging slots are currently bound for
    int64_t ing_ld = 0;
    gat::Resident res;
    bool sample_origin_on = false;   // gat_set_sample_origin: slot sample s is sample origin + s of the period the phases refer to
    int sample_origin = 0;           // (gat_mg sets a signed value: its callers' phases refer to start_sample, not to sample 0)
};

namespace gat {

inline int fail(gat_ctx *ctx, int status, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return status;
}

inline int cuda_fail(gat_ctx *ctx, cudaError_t e, const char *what)
{
    return fail(ctx, GAT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define GAT_CUDA(ctx, call)                                          \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);   \
    } while (0)

inline SignalSlot *find_slot(gat_ctx *ctx, int slot)
{
    auto it = ctx->slots.find(slot);
    return it == ctx->slots.end() ? nullptr : &it->second;
}

inline bool slot_has_signal(const SignalSlot *s) { return s && (s->planes_valid || s->raw_valid || !s->parts.empty()); }

inline int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

inline int pow2_ceil(int x)
{
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

template <typename T>
int ensure_device(gat_ctx *ctx, T *&ptr, size_t &cap, size_t need, bool zero)
{
    if (need <= cap) return GAT_OK;
    // the old buffer may still be in use by work queued on the stream
    GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ptr) GAT_CUDA(ctx, cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    const size_t grow = need + need / 2;
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&ptr), grow * sizeof(T)));
    if (zero) GAT_CUDA(ctx, cudaMemsetAsync(ptr, 0, grow * sizeof(T), ctx->stream));
    cap = grow;
    return GAT_OK;
}

}  // namespace gat
