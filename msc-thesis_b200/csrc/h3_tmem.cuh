// Tensor memory (TMEM) as a per-thread on-chip store for episode state (sm_100a).
//
// A B200 SM has 256 KB of tensor memory next to its 227 KB of shared memory and 256 KB of registers:
// 128 lanes x 512 columns x 32 bit.  The 32x32b shape of tcgen05.ld / tcgen05.st moves, for every thread
// of a warp, N consecutive columns of "its" lane (warp w of a CTA reaches lanes [32 (w % 4), +32), thread
// i lane 32 (w % 4) + i) between TMEM and N registers -- i.e. TMEM is a second thread-private scratchpad
// of 2 KB per thread for a 128-thread CTA.  The closed-loop kernel parks the collective agent's networks
// there (h3_closed_loop.cuh, COLT), which is what lets 8 warps x 32 episodes stay resident per SM in the
// launches where both agents learn.
//
// No tensor-core instruction is involved: the kernel's largest contraction is 9x9.  TMEM is used for its
// capacity.  tcgen05.ld / st are warp-collective (.sync.aligned): every lane of the warp must execute them,
// so callers keep the warp converged around these calls.
#pragma once
#include <cstdint>

namespace h3 {

#define H3_TM_DEV __device__ __forceinline__

// One warp of the CTA allocates `ncols` columns (power of two, 32..512) and publishes the base address
// through shared memory; every thread of the CTA returns it.  Contains __syncthreads().
H3_TM_DEV uint32_t tmem_alloc(uint32_t* sh_slot, int ncols) {
    if ((threadIdx.x >> 5) == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sh_slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *sh_slot;
}

// The allocating warp frees the columns after every warp of the CTA is done with them.  Contains __syncthreads().
H3_TM_DEV void tmem_free(uint32_t base, int ncols) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
    }
}

// Address of this warp's lane quadrant, column `col` (32-bit columns) of an allocation at `base`.
H3_TM_DEV uint32_t tmem_addr(uint32_t base, int col) {
    const uint32_t quad = (threadIdx.x >> 5) & 3u;
    return base + ((quad * 32u) << 16) + (uint32_t)col;
}

H3_TM_DEV void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
H3_TM_DEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- raw 32-bit moves: N registers <-> N columns of this thread's lane --------------------------------
// The loads are asynchronous: the registers are defined only after tmem_wait_ld().  TmemRaw::ld issues
// the wait inside the same asm statement, so the compiler cannot schedule a use in between.
template <int N>
struct TmemRaw;

// (operand lists are spelled out: inline asm has no loops)
template <>
struct TmemRaw<1> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[1]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0])
                     : "r"(a)
                     : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[1]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(a) : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[1]) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(r[0]) : "memory");
    }
};

template <>
struct TmemRaw<2> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[2]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1])
                     : "r"(a)
                     : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[2]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                     : "=r"(r[0]), "=r"(r[1])
                     : "r"(a)
                     : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[2]) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(r[0]), "r"(r[1]) : "memory");
    }
};

template <>
struct TmemRaw<4> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[4]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(a)
                     : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[4]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(a)
                     : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[4]) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r[0]), "r"(r[1]),
                     "r"(r[2]), "r"(r[3])
                     : "memory");
    }
};

template <>
struct TmemRaw<8> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[8]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[8]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[8]) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a), "r"(r[0]),
                     "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                     : "memory");
    }
};

template <>
struct TmemRaw<16> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[16]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15}, [%16];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[16]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[16]) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16};" ::"r"(a),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
            : "memory");
    }
};

template <>
struct TmemRaw<32> {
    H3_TM_DEV static void ld(uint32_t a, uint32_t (&r)[32]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
            "tcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void ld_async(uint32_t a, uint32_t (&r)[32]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(a)
            : "memory");
    }
    H3_TM_DEV static void st(uint32_t a, const uint32_t (&r)[32]) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(a),
            "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
            "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
            "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
            "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
            : "memory");
    }
};

// ---- typed moves: NW elements of T (double = 2 columns, float = 1 column) ----------------------------
// `word` is the element index inside the thread's private column range that starts at `addr`.
template <typename T>
struct TmemCols {
    static constexpr int per = sizeof(T) / 4;
};

// pieces: a move of NW elements is split into power-of-two pieces of at most 32 registers
template <typename T, int NW, int DONE, int SZ>
H3_TM_DEV void tmem_load_piece(uint32_t addr, T (&v)[NW]) {
    constexpr int PER = TmemCols<T>::per;
    constexpr int NR = NW * PER;        // registers
    if constexpr (NR - DONE >= SZ) {
        uint32_t r[SZ];
        TmemRaw<SZ>::ld(addr + (uint32_t)DONE, r);
#pragma unroll
        for (int i = 0; i < SZ / PER; ++i) {
            if constexpr (PER == 2) {
                v[DONE / 2 + i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
            } else {
                v[DONE + i] = __uint_as_float(r[i]);
            }
        }
        tmem_load_piece<T, NW, DONE + SZ, SZ>(addr, v);
    } else if constexpr (SZ > 1) {
        tmem_load_piece<T, NW, DONE, SZ / 2>(addr, v);
    }
}

template <typename T, int NW, int DONE, int SZ>
H3_TM_DEV void tmem_store_piece(uint32_t addr, const T (&v)[NW]) {
    constexpr int PER = TmemCols<T>::per;
    constexpr int NR = NW * PER;
    if constexpr (NR - DONE >= SZ) {
        uint32_t r[SZ];
#pragma unroll
        for (int i = 0; i < SZ / PER; ++i) {
            if constexpr (PER == 2) {
                r[2 * i] = (uint32_t)__double2loint(v[DONE / 2 + i]);
                r[2 * i + 1] = (uint32_t)__double2hiint(v[DONE / 2 + i]);
            } else {
                r[i] = __float_as_uint(v[DONE + i]);
            }
        }
        TmemRaw<SZ>::st(addr + (uint32_t)DONE, r);
        tmem_store_piece<T, NW, DONE + SZ, SZ>(addr, v);
    } else if constexpr (SZ > 1) {
        tmem_store_piece<T, NW, DONE, SZ / 2>(addr, v);
    }
}

// NW elements starting at element `word` of the thread's private range at `addr` (registers valid on return)
template <typename T, int NW>
H3_TM_DEV void tmem_load(uint32_t addr, int word, T (&v)[NW]) {
    tmem_load_piece<T, NW, 0, 32>(addr + (uint32_t)(word * TmemCols<T>::per), v);
}

// asynchronous: tmem_wait_st() before the same columns are loaded again
template <typename T, int NW>
H3_TM_DEV void tmem_store(uint32_t addr, int word, const T (&v)[NW]) {
    tmem_store_piece<T, NW, 0, 32>(addr + (uint32_t)(word * TmemCols<T>::per), v);
}

// ---- asynchronous typed loads: issue(), then ONE tmem_wait_ld(), then read ----------------------------
// The raw registers are converted to T only in operator[] (i.e. after the wait): nothing touches them between
// the tcgen05.ld and the tcgen05.wait::ld.
template <typename T, int NW>
struct TmemBuf {
    static constexpr int PER = TmemCols<T>::per;
    static constexpr int NR = NW * PER;
    uint32_t r[NR];
    template <int DONE, int SZ>
    H3_TM_DEV void piece(uint32_t addr) {
        if constexpr (NR - DONE >= SZ) {
            uint32_t(&dst)[SZ] = *reinterpret_cast<uint32_t(*)[SZ]>(&r[DONE]);
            TmemRaw<SZ>::ld_async(addr + (uint32_t)DONE, dst);
            piece<DONE + SZ, SZ>(addr);
        } else if constexpr (SZ > 1) {
            piece<DONE, SZ / 2>(addr);
        }
    }
    H3_TM_DEV void issue(uint32_t addr, int word) { piece<0, 32>(addr + (uint32_t)(word * PER)); }
    H3_TM_DEV T operator[](int i) const {
        if constexpr (PER == 2) {
            return __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
        } else {
            return __uint_as_float(r[i]);
        }
    }
};

}  // namespace h3
