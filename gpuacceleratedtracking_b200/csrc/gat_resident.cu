// gat_resident.cu -- the correlate kernel as a RESIDENT kernel: launched once, it stays on the device and runs one
// correlation per command the host writes into mapped memory (gat_resident_* in include/gat.h).
//
// Why: the reference times ONE call + synchronisation per 1 ms block (/root/reference/src/benchmarks.jl:872,
// paper/paper.tex:150).  Through a kernel launch such a call costs ~18 us from C on a B200 (4.4 us cooperative launch,
// ~2 us front-end latency, ~6 us kernel with its pipeline fill, ~3 us completion -> host), 0.05 of the 1 us HBM roofline
// of the block.  A resident kernel takes the launch, the argument marshalling and the completion interrupt out of the
// call: host write -> PCIe -> poll -> correlate -> posted result writes + flag -> host poll.
//
// Protocol (one command in flight, sequence numbers 1, 2, ..):
//   host    writes the command (slot index + up to 32 channel records) as 16-byte cells {d0, d1, d2, seq} into pinned mapped
//           memory: the data words first, then the sequence words.  A cell is read by ONE 16-byte load, so a cell whose seq matches carries this
//           command's data whatever order the PCIe reads of different cells complete in.
//   CTA 0   warp 0 polls the cells over PCIe (one warp-wide load per try), then relays them, sequence words included,
//           through device memory; the other CTAs poll those cells in L2 -- 147 CTAs polling host memory would put
//           ~10 GB/s of reads on the PCIe link next to the uploads.
//   all     copy the data words into shared memory (the channel records live there), run correlate_body, meet at the
//           grid barrier, finalise straight into host memory: every accumulator is ONE 8-byte store {value, sequence
//           number}, so the host sees a command complete when all its elements carry the number -- no completion fence
//           (a system-scope fence after the sysmem stores cost 2 - 11 us per CTA), no flag, no done counter.
//   exit    command op 2, or no command for `idle_limit_ms` (then the host relaunches on the next call): a resident
//           kernel owns every SM, so anything else queued on the device waits for it -- the idle limit bounds that wait.
#define GAT_RESIDENT_TU
#include "gat_correlate.cu"

namespace gat {

__device__ __forceinline__ uint4 ld_sys_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <int A, int L, bool HELP, int ROLE>
__device__ __forceinline__ void resident_loop(const CorrArgs &args, const ResCtl &ctl)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint32_t *cmd_s = reinterpret_cast<uint32_t *>(smem + ctl.cmd_off);   // the command's data words, cell by cell
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t seq = ctl.first_seq;; ++seq) {
        if (warp == 0) {
            // a command is n_cells cells = up to kResChunks warp-wide 16-byte loads, all in flight together (one PCIe round trip)
            uint4 c[kResChunks];
            const int n_cells = ctl.n_cells;
            auto load_all = [&](const uint4 *base) {
                bool ok = true;
#pragma unroll
                for (int k = 0; k < kResChunks; ++k) {
                    c[k] = make_uint4(0u, 0u, 0u, seq);
                    if (32 * k + lane < n_cells) c[k] = ld_sys_v4(base + 32 * k + lane);
                }
#pragma unroll
                for (int k = 0; k < kResChunks; ++k) ok = ok && c[k].w == seq;
                return __all_sync(0xffffffffu, ok);
            };
            if (blockIdx.x == 0) {
                const uint64_t t0 = global_timer_ns();
                bool idle = false;
                while (!load_all(ctl.cmd_host)) {
                    if (global_timer_ns() - t0 > (uint64_t)ctl.idle_limit_ms * 1000000ull) {
                        idle = true;
                        break;
                    }
                }
                if (idle) {
#pragma unroll
                    for (int k = 0; k < kResChunks; ++k) c[k] = make_uint4((k == 0 && lane == 0) ? kResOpExit : 0u, 0u, 0u, seq);
                }
                if (ctl.stamps && lane == 0) ctl.stamps[0] = global_timer_ns();     // debug: command seen
                // relay: the same cells (each one 16-byte store, its sequence word included) through device memory
#pragma unroll
                for (int k = 0; k < kResChunks; ++k)
                    if (32 * k + lane < n_cells)
                        asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ctl.relay + 32 * k + lane), "r"(c[k].x), "r"(c[k].y),
                                     "r"(c[k].z), "r"(c[k].w)
                                     : "memory");
            } else {
                unsigned int tries = 0;
                while (!load_all(ctl.relay))
                    if (++tries > 4096u) __nanosleep(200);   // long idle: stop hammering L2
            }
#pragma unroll
            for (int k = 0; k < kResChunks; ++k)
                if (32 * k + lane < n_cells) {
                    cmd_s[3 * (32 * k + lane) + 0] = c[k].x;
                    cmd_s[3 * (32 * k + lane) + 1] = c[k].y;
                    cmd_s[3 * (32 * k + lane) + 2] = c[k].z;
                }
        }
        cta_sync<ROLE>();
        if (cmd_s[0] != kResOpCorrelate) return;    // exit command, or CTA 0 gave up waiting
        if (threadIdx.x == 32) {
            // the selected block's TMA descriptors on their way into the descriptor cache while the barriers are set up
            const PeriodDev *pd = ctl.slot_maps + cmd_s[1];
            asm volatile("prefetch.tensormap [%0];" ::"l"(&pd->re) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&pd->im) : "memory");
        }
        ResOverride ro;
        ro.periods = ctl.slot_maps + cmd_s[1];
        ro.sats = reinterpret_cast<const SatDev *>(cmd_s + 4);
        ro.barrier_target = args.barrier_target + (seq - ctl.first_seq) * gridDim.x;
        ro.seq = seq;
        ro.reinit = seq != ctl.first_seq;
        if (ctl.stamps && blockIdx.x == 0 && threadIdx.x == 0) ctl.stamps[1] = global_timer_ns();   // debug: body entered
        correlate_body<A, L, false, false, false, HELP, ROLE, true>(args, ro);
        if (ctl.stamps && blockIdx.x == 0 && threadIdx.x == 0) ctl.stamps[2] = global_timer_ns();   // debug: CTA 0 through its finalize
        cta_sync<ROLE>();                             // every warp is out of the body before its barriers are set up again
    }
}

template <int A, int L, bool HELP>
__global__ void __launch_bounds__(HELP ? block_threads_help(A, L) : block_threads_max(A, L), 1)
    resident_kernel(const __grid_constant__ CorrArgs args, const __grid_constant__ ResCtl ctl)
{
    if constexpr (HELP && help_realloc(A, L)) {
        if ((threadIdx.x >> 5) >= kReallocConsumerWarps) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kReallocAuxRegs));
            resident_loop<A, L, HELP, kRoleAux>(args, ctl);
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kReallocConsumerRegs));
            resident_loop<A, L, HELP, kRoleConsumer>(args, ctl);
        }
    } else {
        resident_loop<A, L, HELP, kRoleAll>(args, ctl);
    }
}

typedef void (*ResidentFn)(const CorrArgs, const ResCtl);

// the shapes of the reference's one-call-per-block sweep (scripts/run_benchmarks_gpsl1.jl:5-18: 1 / 4 / 16 antennas, 3 / 7
// taps) and the 11-tap monitor shape
static ResidentFn pick_resident(int A, int L, bool help)
{
#define GAT_RES_CASE(a, l, h) \
    if (A == a && L == l && help == h) return (ResidentFn)resident_kernel<a, l, h>;
    GAT_RES_CASE(1, 3, false) GAT_RES_CASE(4, 3, false) GAT_RES_CASE(16, 3, false)
    GAT_RES_CASE(1, 7, false) GAT_RES_CASE(4, 7, false) GAT_RES_CASE(4, 7, true)
    GAT_RES_CASE(4, 11, false) GAT_RES_CASE(4, 11, true)
#undef GAT_RES_CASE
    return nullptr;
}

bool resident_kernel_available(int A, int L, bool help) { return pick_resident(A, L, help) != nullptr; }

cudaError_t launch_resident(const LaunchPlan &plan, const CorrArgs &args, const ResCtl &ctl, size_t smem_bytes, cudaStream_t stream)
{
    ResidentFn fn = pick_resident(plan.A, plan.L, plan.help);
    if (!fn) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    // cooperative: every CTA must be resident for the grid barrier inside each command
    void *kargs[] = {const_cast<CorrArgs *>(&args), const_cast<ResCtl *>(&ctl)};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(fn), dim3(plan.grid), dim3(plan.block), kargs, smem_bytes, stream);
}

}  // namespace gat
