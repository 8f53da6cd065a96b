// gat_ring.cu -- the signal ring: signal blocks whose samples are spread over the GPUs of one box, read by every
// rank's correlate kernel straight out of the owners' HBM (TMA loads over NVLink 5 / NVSwitch peer mappings).
//
// north_star / SURVEY.md 8(e): satellite channels are independent given the signal block, so they shard across GPUs and
// every GPU needs every block.  The reference has no multi-GPU code at all (single CuDevice(0), src/benchmarks.jl:24;
// its only hook is `sat_idx = blockIdx().z`, src/algorithms.jl:656, paper/paper.tex:114).  Instead of a broadcast
// collective followed by the kernel (an extra launch, a receive buffer, one more HBM write + read of every byte, and the
// root's egress as the bottleneck), the exchange is the kernel's own tile pipeline:
//   * a block is cut into `world` contiguous sample ranges of whole 256-sample tiles; rank r owns range r.  The host
//     feeds each range through the owner's PCIe link (gat_ring_upload*: world links in parallel);
//   * every rank maps all owners' allocations once (CUDA IPC across processes, direct peer access inside one process)
//     and builds one pair of TMA descriptors per (slot, part);
//   * the producer warp of the correlate kernel picks the descriptor of the part a tile lives in -- the all-gather
//     happens tile by tile inside the kernel, overlapped with the math of the previous tiles, each byte crossing
//     NVLink exactly once per reader and never touching the reader's HBM.
// Cross-rank ordering uses sequence flags in the ring allocations (peer stores + acquire spins in one-warp kernels),
// stream-ordered on both sides, so the host never blocks:
//   copy stream : acquire(gen - depth)  ->  upload parts of generation gen  ->  publish
//   main stream : wait(gen)             ->  correlate calls reading gen     ->  release
#include "gat_ctx.h"

using namespace gat;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn ring_encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int ring_check(gat_ctx *ctx, bool need_connected)
{
    if (!ctx) return GAT_ERR_INVALID;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    if (!ctx->ring.local) return fail(ctx, GAT_ERR_INVALID, "no signal ring (gat_ring_create first)");
    if (need_connected && !ctx->ring.connected) return fail(ctx, GAT_ERR_INVALID, "signal ring not connected");
    return GAT_OK;
}

// every rank derives the same geometry from (n_samples, world)
void ring_geometry(Ring &r)
{
    const int tiles = (r.n_samples + kTileCap - 1) / kTileCap;
    r.part_tiles = (tiles + r.world - 1) / r.world;
    r.n_parts = (tiles + r.part_tiles - 1) / r.part_tiles;
    for (int j = 0; j < r.world; ++j) {
        const int lo = std::min(r.n_samples, j * r.part_tiles * kTileCap);
        const int hi = std::min(r.n_samples, (j + 1) * r.part_tiles * kTileCap);
        r.part_start[j] = lo;
        r.part_len[j] = hi - lo;
        r.part_ld[j] = std::max<int64_t>(4, (static_cast<int64_t>(hi - lo) + 3) & ~3LL);
    }
}

size_t ring_slot_bytes(const Ring &r, int j) { return 2 * static_cast<size_t>(r.n_ants) * r.part_ld[j] * sizeof(float); }

float *ring_plane(const Ring &r, int j, int slot, int plane)
{
    return reinterpret_cast<float *>(r.base[j] + kRingFlagBytes + static_cast<size_t>(slot) * ring_slot_bytes(r, j)) +
           static_cast<size_t>(plane) * r.n_ants * r.part_ld[j];
}

unsigned int *ring_flags(const Ring &r, int j, int which) { return reinterpret_cast<unsigned int *>(r.base[j]) + 16 * which; }

float *view_plane(const Ring &r, unsigned char *const *bases, int j, int slot, int plane)
{
    return reinterpret_cast<float *>(bases[j] + kRingFlagBytes + static_cast<size_t>(slot) * ring_slot_bytes(r, j)) +
           static_cast<size_t>(plane) * r.n_ants * r.part_ld[j];
}

// turn ctx slots slot_base .. slot_base + n_slots-1 into sharded slots whose part j lives at bases[j]
int ring_build_view(gat_ctx *ctx, int slot_base, unsigned char *const *bases)
{
    Ring &r = ctx->ring;
    EncodeTiledFn enc = ring_encode_fn();
    if (!enc) return fail(ctx, GAT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    for (int s = 0; s < r.n_slots; ++s) {
        SignalSlot &sl = ctx->slots[slot_base + s];
        if ((sl.owned && sl.re) || sl.raw || sl.peer_base) {
            GAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (sl.owned && sl.re) GAT_CUDA(ctx, cudaFree(sl.re));
            if (sl.raw) GAT_CUDA(ctx, cudaFree(sl.raw));
            if (sl.peer_base) GAT_CUDA(ctx, cudaIpcCloseMemHandle(sl.peer_base));
        }
        sl = SignalSlot{};
        sl.n_samples = r.n_samples;
        sl.n_ants = r.n_ants;
        sl.part_tiles = r.part_tiles;
        sl.parts.resize(r.n_parts);
        for (int j = 0; j < r.n_parts; ++j) {
            SlotPart &pt = sl.parts[j];
            pt.re = view_plane(r, bases, j, s, 0);
            pt.im = view_plane(r, bases, j, s, 1);
            pt.ld = r.part_ld[j];
            pt.start = r.part_start[j];
            pt.len = r.part_len[j];
            const cuuint64_t dims[2] = {static_cast<cuuint64_t>(pt.len), static_cast<cuuint64_t>(r.n_ants)};
            const cuuint64_t strides[1] = {static_cast<cuuint64_t>(pt.ld) * sizeof(float)};
            const cuuint32_t box[2] = {static_cast<cuuint32_t>(kTileCap), static_cast<cuuint32_t>(r.n_ants)};
            const cuuint32_t estr[2] = {1, 1};
            float *planes[2] = {pt.re, pt.im};
            CUtensorMap *maps[2] = {&pt.maps.re, &pt.maps.im};
            for (int i = 0; i < 2; ++i) {
                CUresult cr = enc(maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, planes[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (cr != CUDA_SUCCESS)
                    return fail(ctx, GAT_ERR_ALIGNMENT, "cuTensorMapEncodeTiled rejected a ring part (CUresult " + std::to_string(cr) + ")");
            }
        }
    }
    return GAT_OK;
}

// all mappings are in place: ctx slots 0 .. n_slots-1 become the ring's blocks (parts read where they live)
int ring_build_slots(gat_ctx *ctx)
{
    int rc = ring_build_view(ctx, 0, ctx->ring.base);
    if (rc) return rc;
    ctx->ring.connected = true;
    return GAT_OK;
}

int ring_signal(gat_ctx *ctx, int which, unsigned int seq, cudaStream_t stream)
{
    Ring &r = ctx->ring;
    FlagPtrs dst{};
    for (int j = 0; j < r.world; ++j) dst.p[j] = ring_flags(r, j, which);
    cudaError_t e = launch_flag_signal(dst, r.world, r.rank, seq, stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "ring flag launch");
    ctx->launches += 1;
    return GAT_OK;
}

int ring_wait_flags(gat_ctx *ctx, int which, int seq, cudaStream_t stream)
{
    if (seq <= 0) return GAT_OK;
    Ring &r = ctx->ring;
    cudaError_t e = launch_flag_wait(ring_flags(r, r.rank, which), r.world, static_cast<unsigned int>(seq), stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "ring wait launch");
    ctx->launches += 1;
    return GAT_OK;
}

}  // namespace

extern "C" {

int gat_ring_destroy(gat_ctx *ctx)
{
    if (!ctx) return GAT_ERR_INVALID;
    cudaSetDevice(ctx->device);
    Ring &r = ctx->ring;
    if (!r.local) return GAT_OK;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int s = 0; s < 2 * r.n_slots; ++s) {
        auto it = ctx->slots.find(s);
        if (it != ctx->slots.end() && !it->second.parts.empty()) ctx->slots.erase(it);
    }
    for (int j = 0; j < kMaxPeers; ++j) {
        if (r.opened[j]) cudaIpcCloseMemHandle(r.opened[j]);
        if (r.mirror[j] && r.mirror[j] != r.local) cudaFree(r.mirror[j]);
    }
    for (int i = 0; i < kRingEvents; ++i) {
        if (r.rel_ev[i]) cudaEventDestroy(r.rel_ev[i]);
        if (r.pf_ev[i]) cudaEventDestroy(r.pf_ev[i]);
    }
    cudaFree(r.local);
    r = Ring{};
    return GAT_OK;
}

int gat_ring_create(gat_ctx *ctx, int world, int rank, int n_slots, int n_samples, int n_ants, unsigned char *handle_out)
{
    if (!ctx) return GAT_ERR_INVALID;
    GAT_CUDA(ctx, cudaSetDevice(ctx->device));
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_slots < 1 || n_slots > 32768 || n_samples < 1 || n_ants < 1 ||
        n_ants > kMaxAnts || !handle_out)
        return fail(ctx, GAT_ERR_INVALID, "bad ring arguments (1 <= world <= 8, 1 <= n_slots <= 32768)");
    int rc = gat_ring_destroy(ctx);
    if (rc) return rc;
    Ring &r = ctx->ring;
    r.world = world;
    r.rank = rank;
    r.n_slots = n_slots;
    r.n_samples = n_samples;
    r.n_ants = n_ants;
    ring_geometry(r);
    const size_t bytes = kRingFlagBytes + static_cast<size_t>(n_slots) * ring_slot_bytes(r, rank);
    GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&r.local), bytes));
    GAT_CUDA(ctx, cudaMemset(r.local, 0, bytes));
    r.base[rank] = r.local;
    cudaIpcMemHandle_t h;
    GAT_CUDA(ctx, cudaIpcGetMemHandle(&h, r.local));
    static_assert(sizeof(h) == GAT_IPC_HANDLE_BYTES, "IPC handle size");
    std::memcpy(handle_out, &h, sizeof(h));
    return GAT_OK;
}

int gat_ring_connect(gat_ctx *ctx, const unsigned char *handles)
{
    int rc = ring_check(ctx, false);
    if (rc) return rc;
    if (!handles) return fail(ctx, GAT_ERR_INVALID, "null handles");
    Ring &r = ctx->ring;
    for (int j = 0; j < r.world; ++j) {
        if (j == r.rank || r.base[j]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + static_cast<size_t>(j) * GAT_IPC_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        GAT_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        r.opened[j] = p;
        r.base[j] = static_cast<unsigned char *>(p);
    }
    return ring_build_slots(ctx);
}

int gat_ring_connect_local(gat_ctx *ctx, gat_ctx *const *peers)
{
    int rc = ring_check(ctx, false);
    if (rc) return rc;
    if (!peers) return fail(ctx, GAT_ERR_INVALID, "null peers");
    Ring &r = ctx->ring;
    for (int j = 0; j < r.world; ++j) {
        if (j == r.rank) continue;
        gat_ctx *pc = peers[j];
        if (!pc || !pc->ring.local || pc->ring.world != r.world || pc->ring.rank != j || pc->ring.n_slots != r.n_slots ||
            pc->ring.n_samples != r.n_samples || pc->ring.n_ants != r.n_ants)
            return fail(ctx, GAT_ERR_INVALID, "peer " + std::to_string(j) + " has no matching ring");
        if (pc->device != ctx->device) {
            int can = 0;
            GAT_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, pc->device));
            if (!can) return fail(ctx, GAT_ERR_UNSUPPORTED, "no peer access between devices " + std::to_string(ctx->device) + " and " + std::to_string(pc->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(pc->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
        }
        r.base[j] = pc->ring.local;
    }
    return ring_build_slots(ctx);
}

int gat_ring_part(gat_ctx *ctx, int rank, int *start_out, int *len_out)
{
    if (!ctx || !ctx->ring.local) return GAT_ERR_INVALID;
    if (rank < 0 || rank >= ctx->ring.world) return fail(ctx, GAT_ERR_INVALID, "rank out of range");
    if (start_out) *start_out = ctx->ring.part_start[rank];
    if (len_out) *len_out = ctx->ring.part_len[rank];
    return GAT_OK;
}

int gat_ring_upload_part(gat_ctx *ctx, int slot, const float *re_part, const float *im_part, int ld, int src_is_device)
{
    NvtxRange nvtx_call("gat_ring_upload");
    int rc = ring_check(ctx, false);
    if (rc) return rc;
    Ring &r = ctx->ring;
    if (slot < 0 || slot >= r.n_slots) return fail(ctx, GAT_ERR_INVALID, "ring slot out of range");
    const int len = r.part_len[r.rank];
    if (len == 0) return GAT_OK;                      // a short block: this rank owns no samples of it
    if (!re_part || !im_part || ld < len) return fail(ctx, GAT_ERR_INVALID, "bad ring upload arguments");
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t dpitch = static_cast<size_t>(r.part_ld[r.rank]) * sizeof(float), spitch = static_cast<size_t>(ld) * sizeof(float);
    const size_t width = static_cast<size_t>(len) * sizeof(float);
    float *d_re = ring_plane(r, r.rank, slot, 0), *d_im = ring_plane(r, r.rank, slot, 1);
    if (dpitch == spitch) {
        // same pitch on both sides: one flat copy per plane (the last row's padding is not read)
        const size_t bytes = spitch * (r.n_ants - 1) + width;
        GAT_CUDA(ctx, cudaMemcpyAsync(d_re, re_part, bytes, kind, ctx->copy_stream));
        GAT_CUDA(ctx, cudaMemcpyAsync(d_im, im_part, bytes, kind, ctx->copy_stream));
    } else {
        GAT_CUDA(ctx, cudaMemcpy2DAsync(d_re, dpitch, re_part, spitch, width, r.n_ants, kind, ctx->copy_stream));
        GAT_CUDA(ctx, cudaMemcpy2DAsync(d_im, dpitch, im_part, spitch, width, r.n_ants, kind, ctx->copy_stream));
    }
    return GAT_OK;
}

int gat_ring_upload(gat_ctx *ctx, int slot, const float *re, const float *im, int ld, int src_is_device)
{
    if (!ctx || !ctx->ring.local) return GAT_ERR_INVALID;
    if (!re || !im || ld < ctx->ring.n_samples) return fail(ctx, GAT_ERR_INVALID, "bad ring upload arguments");
    const int start = ctx->ring.part_start[ctx->ring.rank];
    return gat_ring_upload_part(ctx, slot, re + start, im + start, ld, src_is_device);
}

int gat_ring_publish(gat_ctx *ctx)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    rc = ring_signal(ctx, 0, ctx->ring.pub_seq + 1, ctx->copy_stream);
    if (rc) return rc;
    return static_cast<int>(++ctx->ring.pub_seq);
}

int gat_ring_wait(gat_ctx *ctx, int generation)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    return ring_wait_flags(ctx, 0, generation, ctx->stream);
}

int gat_ring_release(gat_ctx *ctx)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    rc = ring_signal(ctx, 1, ctx->ring.rel_seq + 1, ctx->stream);
    if (rc) return rc;
    ++ctx->ring.rel_seq;
    if (ctx->ring.mirrored) GAT_CUDA(ctx, cudaEventRecord(ctx->ring.rel_ev[ctx->ring.rel_seq % kRingEvents], ctx->stream));
    return static_cast<int>(ctx->ring.rel_seq);
}

// ---- mirror view: for shapes whose kernel is compute-bound (many satellites per GPU and block) ----
// The fused pull fetches a remote tile once per satellite GROUP of the reading CTA set (remote lines are not cached in
// the reader's L2), and its NVLink time sits on the kernel's critical path.  With the mirror the copy engines bring the
// peers' shares into local HBM one generation ahead (no SM involved, fully under the previous generation's kernel);
// slots n_slots .. 2 n_slots-1 are the same blocks read from those local copies.
int gat_ring_enable_mirror(gat_ctx *ctx)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    Ring &r = ctx->ring;
    if (r.mirrored) return GAT_OK;
    for (int j = 0; j < r.n_parts; ++j) {
        if (j == r.rank) {
            r.mirror[j] = r.local;
            continue;
        }
        const size_t bytes = kRingFlagBytes + static_cast<size_t>(r.n_slots) * ring_slot_bytes(r, j);
        GAT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&r.mirror[j]), bytes));
        GAT_CUDA(ctx, cudaMemset(r.mirror[j], 0, bytes));
    }
    for (int i = 0; i < kRingEvents; ++i) {
        GAT_CUDA(ctx, cudaEventCreateWithFlags(&r.rel_ev[i], cudaEventDisableTiming));
        GAT_CUDA(ctx, cudaEventCreateWithFlags(&r.pf_ev[i], cudaEventDisableTiming));
    }
    rc = ring_build_view(ctx, r.n_slots, r.mirror);
    if (rc) return rc;
    r.mirrored = true;
    return GAT_OK;
}

int gat_ring_prefetch(gat_ctx *ctx, int first_slot, int n_slots, int generation, int releases)
{
    NvtxRange nvtx_call("gat_ring_prefetch");
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    Ring &r = ctx->ring;
    if (!r.mirrored) return fail(ctx, GAT_ERR_INVALID, "gat_ring_enable_mirror first");
    if (first_slot < 0 || n_slots < 1 || first_slot + n_slots > r.n_slots) return fail(ctx, GAT_ERR_INVALID, "ring slot range out of bounds");
    rc = ring_wait_flags(ctx, 0, generation, ctx->copy_stream);         // every owner's share of that generation is in place
    if (rc) return rc;
    if (releases > 0 && static_cast<unsigned int>(releases) <= r.rel_seq)     // this rank's own kernels are done with the mirror slots
        GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, r.rel_ev[releases % kRingEvents], 0));
    for (int j = 0; j < r.n_parts; ++j) {
        if (j == r.rank || r.part_len[j] == 0) continue;
        const size_t off = kRingFlagBytes + static_cast<size_t>(first_slot) * ring_slot_bytes(r, j);
        GAT_CUDA(ctx, cudaMemcpyAsync(r.mirror[j] + off, r.base[j] + off, static_cast<size_t>(n_slots) * ring_slot_bytes(r, j),
                                      cudaMemcpyDeviceToDevice, ctx->copy_stream));
    }
    const int ticket = ++r.pf_tickets;
    GAT_CUDA(ctx, cudaEventRecord(r.pf_ev[ticket % kRingEvents], ctx->copy_stream));
    return ticket;
}

int gat_ring_mirror_wait(gat_ctx *ctx, int ticket)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    Ring &r = ctx->ring;
    if (!r.mirrored || ticket < 1 || ticket > r.pf_tickets) return fail(ctx, GAT_ERR_INVALID, "no such prefetch ticket");
    if (r.pf_tickets - ticket >= kRingEvents) return GAT_OK;            // long done: its event has been re-used since
    GAT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, r.pf_ev[ticket % kRingEvents], 0));
    return GAT_OK;
}

int gat_ring_acquire(gat_ctx *ctx, int releases)
{
    int rc = ring_check(ctx, true);
    if (rc) return rc;
    return ring_wait_flags(ctx, 1, releases, ctx->copy_stream);
}

}  // extern "C"
