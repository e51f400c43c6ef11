// Host-side dispatch of the fused closed-loop kernel; one translation unit per precision pair
// instantiates launch_closed_loop<ET, NT> (h3_loop_*.cu) so the variants compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "h3_closed_loop.cuh"
#include "h3_closed_loop_tm.cuh"
#include "h3_closed_loop_rt.cuh"

namespace h3 {

void set_error(const char* fmt, ...);

struct LoopInfo {
    int block, smem_bytes, blocks_per_sm, regs, lanes_per_warp;
};

template <typename ET, typename NT>
int launch_closed_loop(const H3Config& cfg, const H3Batch& b, int step0, const double* ref_table,
                       const double* excitation, int exc_len, int n_ticks, const H3Traj* traj, double* stats,
                       void* stats_scratch, cudaStream_t stream, int force_phase2, LoopInfo* info_only);

#ifdef H3_LOOP_IMPL

template <typename ET, typename NT>
struct KernelSel {
    void (*fn)(const LoopParams<ET, NT>);
    int block, lanes, warps;
    size_t smem;
    int tm_cols = 0;        // tensor-memory columns a CTA allocates (tensor-memory kernel), 0: none
};

// Kernel variants (a pruned cross product, see closed_loop_kernel's template flags):
//   FAST    : symmetric covariance, launch-uniform hyper-parameters, tau_lon == 1   (main_3dof.py config)
//   SWEEP   : symmetric covariance, per-episode (or tau_lon != 1) hyper-parameters  (main.py-style sweeps)
//   GENERAL : covariance not known symmetric (model.py:53-54 literally), any hyper-parameters
enum { VAR_FAST = 0, VAR_SWEEP = 1, VAR_GENERAL = 2 };

constexpr size_t SMEM_PER_BLOCK_MAX = 232448;   // 227 KB opt-in limit per CTA on sm_100
constexpr size_t SMEM_PER_SM = 233472;          // 228 KB per SM, 1 KB of it reserved per resident CTA
// static shared memory of the fused kernels: the tanh table of h3_math.cuh (fp64 networks) + the epilogue's flag
constexpr size_t SMEM_STATIC = ((sizeof(double) * H3_TANHT_N + 127) / 128) * 128 + 128;

// On-chip placement per variant.  Phase-1 launches (COL == false) keep everything
// but the RLS matrices in registers and run two 4-warp CTAs per SM (the CTA is the unit that stages the tanh table).
// Launches with the collective agent go to the tensor-memory kernel (below) when they are lockstep and the covariance
// is symmetric; the rest (heterogeneous batches, the literal non-symmetric update) run one 4-warp CTA per SM with as
// many episodes per warp as the shared-memory tile allows.
// ---- tuning knobs (compile-time; defaults are the measured best) ----
#ifndef H3_SWEEP_P1_REGS
#define H3_SWEEP_P1_REGS 1
#endif
#ifndef H3_RD_P1
#define H3_RD_P1 1
#endif
#ifndef H3_RD_P2
#define H3_RD_P2 2
#endif

template <typename ET, typename NT, int NHL, int NHC, bool COL, int VAR>
struct Tune {
    static constexpr bool SYM = VAR != VAR_GENERAL, HYP = VAR != VAR_FAST, TAU1L = VAR == VAR_FAST;
    // the full cyclic agent (11 nh words) goes to the tile in collective-agent launches of the SWEEP / GENERAL
    // variants; phase-1 launches keep it in registers (7 instead of 4 warps per SM)
    static constexpr bool lon_smem = !TAU1L && (COL || !H3_SWEEP_P1_REGS);
    static constexpr int MINB = COL ? 1 : 2;
    static constexpr int PLACE = (COL ? P_STASHC : 0) | (lon_smem ? (P_LON | P_STASHL) : 0);
    static constexpr int UL = lon_smem ? NHL / 2 : NHL;
    static constexpr int UC = COL ? NHC / 2 : NHC;
    // four warps per CTA, two CTAs per SM: the tanh table is staged once per CTA
    static constexpr int WARPS = 4;
    static constexpr int CTAS = COL ? 1 : 2;
    static constexpr int RD = COL ? H3_RD_P2 : H3_RD_P1;    // RLS stream look-ahead (groups)
    typedef LoopLayout<ET, NT, NHL, NHC, COL, SYM, HYP, TAU1L, PLACE, 32> L32;
    static constexpr size_t room = ((COL ? SMEM_PER_BLOCK_MAX : SMEM_PER_SM / CTAS - 1024) - SMEM_STATIC - L32::head_bytes(WARPS)) / WARPS - 16;
    static constexpr int fit = (int)(room / L32::bytes_per_lane);
    static constexpr int LANES = fit >= 32 ? 32 : fit;
    static_assert(LANES >= 8, "episode state does not fit the shared-memory tile");
};

template <typename ET, typename NT, int NHL, int NHC, bool COL, int VAR>
static KernelSel<ET, NT> sel_kernel() {
    typedef Tune<ET, NT, NHL, NHC, COL, VAR> T;
    typedef LoopLayout<ET, NT, NHL, NHC, COL, T::SYM, T::HYP, T::TAU1L, T::PLACE, T::LANES> LY;
    KernelSel<ET, NT> k;
    k.fn = closed_loop_kernel<ET, NT, NHL, NHC, COL, T::SYM, T::HYP, T::TAU1L, T::PLACE, T::LANES, T::WARPS, T::UL,
                              T::UC, T::RD, T::MINB>;
    k.block = T::WARPS * 32;
    k.lanes = T::LANES;
    k.warps = T::WARPS;
    k.smem = LY::smem_bytes(T::WARPS);
    return k;
}

template <typename ET, typename NT, int NHL, int NHC>
static KernelSel<ET, NT> sel_flags(bool col, int variant) {
    if (col) {
        if (variant == VAR_FAST) return sel_kernel<ET, NT, NHL, NHC, true, VAR_FAST>();
        if (variant == VAR_SWEEP) return sel_kernel<ET, NT, NHL, NHC, true, VAR_SWEEP>();
        return sel_kernel<ET, NT, NHL, NHC, true, VAR_GENERAL>();
    }
    if (variant == VAR_FAST) return sel_kernel<ET, NT, NHL, NHC, false, VAR_FAST>();
    if (variant == VAR_SWEEP) return sel_kernel<ET, NT, NHL, NHC, false, VAR_SWEEP>();
    return sel_kernel<ET, NT, NHL, NHC, false, VAR_GENERAL>();
}


// ---- launches with the collective agent, lockstep: the tensor-memory kernel (h3_closed_loop_tm.cuh) ----
#ifndef H3_TM
#define H3_TM 1             // 0: keep the shared-memory-only variants (PAIR / 4-warp) for these launches
#endif
#ifndef H3_TM_WARPS
#define H3_TM_WARPS 4       // warps per CTA of the tensor-memory kernel (8 / H3_TM_WARPS CTAs per SM)
#endif
#ifndef H3_TM_V2
#define H3_TM_V2 2          // ... of its weight update (pass 2)
#endif
#ifndef H3_TM_RD
#define H3_TM_RD 1          // look-ahead (groups) of the RLS shared-memory streams in the tensor-memory kernel
#endif
#ifndef H3_TM_UP
#define H3_TM_UP 5          // hidden units of the collective agent whose critic + target tanh chains are interleaved
#endif

#ifndef H3_TM_P1_SWEEP
#define H3_TM_P1_SWEEP 1    // per-episode-hyper launches that end before the switch: tensor-memory kernel with the cyclic
                            // agent's target critic in TMEM (0: shared-memory kernel, target in registers + spills)
#endif

template <typename ET, typename NT, int NHL, int NHC, int VAR, bool COLT = true>
struct TuneTm {
    static constexpr bool HYP = VAR != VAR_FAST, TAU1L = VAR == VAR_FAST;
    typedef TmLayout<ET, NT, NHL, NHC, HYP, TAU1L, 32, H3_TM_WARPS, COLT> L32;
    // eight warps per SM in one or two CTAs: the tile of a warp gets an eighth of what is left of the SM's shared memory
    static constexpr int CTAS = 8 / H3_TM_WARPS;
    static constexpr size_t room = (SMEM_PER_SM / CTAS - 1024 - SMEM_STATIC - L32::head_bytes) / L32::WARPS - 16;
    static constexpr int fit = (int)(room / L32::bytes_per_col);       // columns incl. the idle lanes' dummy column
    static constexpr int LANES = fit >= 32 ? 32 : fit - 1;
    static_assert(LANES >= 8, "episode state does not fit the shared-memory tile");
    static constexpr int UF = (NHC % 5 == 0) ? 5 : ((NHC % 4 == 0) ? 4 : ((NHC % 3 == 0) ? 3 : ((NHC % 2 == 0) ? 2 : 1)));
    static constexpr int UP = (NHC % H3_TM_UP == 0) ? H3_TM_UP : ((NHC % 4 == 0) ? 4 : ((NHC % 3 == 0) ? 3 : ((NHC % 2 == 0) ? 2 : 1)));
    static constexpr int V2 = (NHC % H3_TM_V2 == 0) ? H3_TM_V2 : 1;
    typedef TmLayout<ET, NT, NHL, NHC, HYP, TAU1L, LANES, H3_TM_WARPS, COLT> LY;
};

template <typename ET, typename NT, int NHL, int NHC, int VAR, bool COLT = true>
static KernelSel<ET, NT> sel_tm() {
    typedef TuneTm<ET, NT, NHL, NHC, VAR, COLT> T;
    KernelSel<ET, NT> k;
    k.fn = closed_loop_tm_kernel<ET, NT, NHL, NHC, T::HYP, T::TAU1L, T::LANES, H3_TM_RD, T::UF, T::UP, T::V2, H3_TM_WARPS, COLT>;
    k.block = T::LY::WARPS * 32;
    k.lanes = T::LANES;
    k.warps = T::LY::WARPS;
    k.smem = T::LY::smem_bytes;
    k.tm_cols = T::LY::tm_cols * (T::LY::WARPS / 4);
    return k;
}

template <typename ET, typename NT, int NHL, int NHC>
static KernelSel<ET, NT> sel_flags(bool col, int variant, bool tm) {
    if (col && tm && variant == VAR_FAST) return sel_tm<ET, NT, NHL, NHC, VAR_FAST>();
    if (col && tm && variant == VAR_SWEEP) return sel_tm<ET, NT, NHL, NHC, VAR_SWEEP>();
    if (!col && tm && variant == VAR_SWEEP) return sel_tm<ET, NT, NHL, NHC, VAR_SWEEP, false>();
    return sel_flags<ET, NT, NHL, NHC>(col, variant);
}

template <typename ET, typename NT>
int launch_closed_loop(const H3Config& cfg, const H3Batch& b, int step0, const double* ref_table,
                       const double* excitation, int exc_len, int n_ticks, const H3Traj* traj, double* stats,
                       void* stats_scratch, cudaStream_t stream, int force_phase2, LoopInfo* info_only) {
    const int nhl = cfg.lon.n_hidden, nhc = cfg.col.n_hidden;
    // the collective agent's networks are only touched by ticks >= switch_step (main_3dof.py:127-137)
    bool col = true;
    if (step0 >= 0 && (int64_t)step0 + n_ticks <= (int64_t)cfg.switch_step) col = false;
    if (force_phase2 >= 0) col = force_phase2 != 0;
    const bool per_ep = b.col_hyper != nullptr || b.lon_hyper != nullptr;
    const int variant = (cfg.rls_symmetric == 0) ? VAR_GENERAL
                                                 : ((per_ep || cfg.lon.tau != 1.0) ? VAR_SWEEP : VAR_FAST);
    // tensor-memory kernel: needs a warp-uniform tick index (lockstep) and the symmetric covariance
    const bool tm = H3_TM && (col || (H3_TM_P1_SWEEP && variant == VAR_SWEEP)) && (step0 >= 0 || info_only != nullptr) &&
                    variant != VAR_GENERAL && cfg.store_policy != 3;
    KernelSel<ET, NT> k;
    if (nhl == 6 && nhc == 10) k = sel_flags<ET, NT, 6, 10>(col, variant, tm);
#ifndef H3_FAST_BUILD
    else if (nhl == 8 && nhc == 8) k = sel_flags<ET, NT, 8, 8>(col, variant, tm);
#endif
    else {
        // any other width: the runtime-width kernel (h3_closed_loop_rt.cuh), everything in global memory
        if (nhl < 1 || nhc < 1) {
            set_error("h3_closed_loop: hidden widths must be positive (lon %d, col %d)", nhl, nhc);
            return H3_E_ARG;
        }
        k.fn = nullptr;
        k.block = RT_BLOCK;
        k.lanes = 32;
        k.warps = RT_BLOCK / 32;
        k.smem = 0;
    }
    const bool rt = k.fn == nullptr;
    void (*rt_fn)(const LoopParams<ET, NT>, int, int) = closed_loop_rt_kernel<ET, NT>;
    cudaError_t err = cudaSuccess;
    if (!rt) err = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k.smem);
    if (!rt && err == cudaSuccess)  // every variant is sized to fill the SM's shared memory with co-resident CTAs
        err = cudaFuncSetAttribute(k.fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (err != cudaSuccess) {
        set_error("cudaFuncSetAttribute(smem=%zu): %s", k.smem, cudaGetErrorString(err));
        return H3_E_CUDA;
    }
    if (info_only != nullptr) {
        cudaFuncAttributes fa;
        err = rt ? cudaFuncGetAttributes(&fa, rt_fn) : cudaFuncGetAttributes(&fa, k.fn);
        if (err != cudaSuccess) {
            set_error("cudaFuncGetAttributes: %s", cudaGetErrorString(err));
            return H3_E_CUDA;
        }
        int nb = 0;
        err = rt ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rt_fn, k.block, k.smem)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k.fn, k.block, k.smem);
        if (err != cudaSuccess) {
            set_error("cudaOccupancyMaxActiveBlocksPerMultiprocessor: %s", cudaGetErrorString(err));
            return H3_E_CUDA;
        }
        if (!rt && k.tm_cols > 0) {
            // The occupancy API answers 1 for any kernel that allocates tensor memory (it cannot know how many of the 512
            // columns the kernel asks for at run time).  What the SM really holds: the minimum over shared memory
            // (dynamic + static + 1 KB reserved per CTA), registers (allocated in units of 8 per thread) and columns
            // -- ncu agrees (launch__occupancy_limit_registers / _shared_mem 2, 7.6 warps active per cycle).
            const int by_smem = (int)(SMEM_PER_SM / (k.smem + fa.sharedSizeBytes + 1024));
            const int by_regs = 65536 / (((fa.numRegs + 7) / 8) * 8 * k.block);
            const int by_tmem = 512 / k.tm_cols;
            nb = by_smem < by_regs ? by_smem : by_regs;
            nb = nb < by_tmem ? nb : by_tmem;
        }
        info_only->block = k.block;
        info_only->smem_bytes = (int)k.smem;
        info_only->blocks_per_sm = nb;
        info_only->regs = fa.numRegs;
        info_only->lanes_per_warp = k.lanes;
        return H3_OK;
    }
    if (b.n <= 0 || n_ticks <= 0) return H3_OK;
    LoopParams<ET, NT> P;
    P.K = make_envk<ET>(cfg);
    P.n = b.n;
    P.ld = b.ld;
    P.state = (ET*)b.state;
    P.ref = (ET*)b.ref;
    P.corr_u = (ET*)b.corr_u;
    P.pid_corr = (ET*)b.pid_corr;
    P.rls_W = (ET*)b.rls_W;
    P.rls_cov = (ET*)b.rls_cov;
    P.rls_x = (ET*)b.rls_x;
    P.rls_u = (ET*)b.rls_u;
    P.t = b.t;
    P.acc = b.acc;
    P.trim = b.trim;
    P.theta_trim = b.theta_trim;
    P.ref_par = b.ref_par;
    P.col_hyper = b.col_hyper;
    P.lon_hyper = b.lon_hyper;
    P.rls_skip = b.rls_skip;
    P.step = b.step;
    P.flags = b.flags;
    P.term_step = b.term_step;
    P.heat = b.heat;
    P.col_net = (NT*)b.col_net;
    P.lon_net = (NT*)b.lon_net;
    P.ref_table = ref_table;
    P.excitation = excitation;
    P.exc_len = exc_len;
    P.n_ticks = n_ticks;
    P.step0 = step0;
    P.task = cfg.task;
    P.ref_mode = cfg.ref_mode;
    P.switch_step = cfg.switch_step;
    P.update_lon = cfg.update_lon;
    P.win_lo = cfg.win_lo;
    P.win_hi = cfg.win_hi;
    P.ref_amp = cfg.ref_amp;
    P.ref_period = cfg.ref_period;
    P.dt = cfg.env.dt;
    P.col = cfg.col;
    P.lon = cfg.lon;
    const H3AgentHyper* hs[2] = {&cfg.col, &cfg.lon};
    AgentH<NT>* hd[2] = {&P.col_h, &P.lon_h};
    for (int a = 0; a < 2; ++a) {
        hd[a]->gamma = (NT)hs[a]->gamma;
        hd[a]->lr_a = (NT)hs[a]->lr_actor;
        hd[a]->lr_c = (NT)hs[a]->lr_critic;
        hd[a]->tau = (NT)hs[a]->tau;
        hd[a]->scale = (NT)hs[a]->scale;
    }
    P.col_wr = (ET)cfg.col.reward_weight;
    P.lon_wr = (ET)cfg.lon.reward_weight;
    P.traj = (traj != nullptr) ? (ET*)traj->out : nullptr;
    P.n_traj = (traj != nullptr) ? traj->n_traj : 0;
    P.ld_traj = (traj != nullptr) ? traj->ld_traj : 0;
    const int64_t per_block = (int64_t)k.lanes * k.warps;
    const int64_t grid = (b.n + per_block - 1) / per_block;
    P.stats = stats;
    P.ticket = nullptr;
    P.partials = nullptr;
    if (stats != nullptr) {
        if (stats_scratch == nullptr) {
            set_error("h3_closed_loop: stats requested without scratch");
            return H3_E_ARG;
        }
        P.ticket = (unsigned int*)stats_scratch;
        P.partials = (double*)((char*)stats_scratch + 16);
        err = cudaMemsetAsync(stats_scratch, 0, 16, stream);
        if (err != cudaSuccess) {
            set_error("cudaMemsetAsync: %s", cudaGetErrorString(err));
            return H3_E_CUDA;
        }
    }
    if (rt)
        rt_fn<<<(unsigned)grid, k.block, 0, stream>>>(P, nhl, nhc);
    else
        k.fn<<<(unsigned)grid, k.block, k.smem, stream>>>(P);
    err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("closed_loop_kernel launch (grid %lld, block %d, smem %zu): %s", (long long)grid, k.block, k.smem,
                  cudaGetErrorString(err));
        return H3_E_CUDA;
    }
    return H3_OK;
}

#endif  // H3_LOOP_IMPL

}  // namespace h3
