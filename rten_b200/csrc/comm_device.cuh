// Device side of the cross-rank (min, max) exchange over NVLink peer mailboxes (comm.cu owns the mailboxes; rowops.cu's
// quantise kernels run the exchange as their prologue so that a batch-sharded DynamicQuantizeLinear costs no extra launch).
#pragma once
#include <cstdint>

namespace rtb {

constexpr int MAX_PEERS = 16;

// one rank's mailbox: slot[parity][sender] = {(epoch << 32) | min, (epoch << 32) | max}
struct Mailbox {
    unsigned long long slot[2][MAX_PEERS][2];
    unsigned epoch;     // exchanges completed by the owner (advanced by the exchanging warp)
    unsigned timeouts;  // exchanges that gave up waiting for a peer (reported by the next host call)
    unsigned ready;     // fused form: 1 once block 0 has written the reduced range of the current launch
    unsigned finished;  // fused form: blocks of the current launch that are done (the last one re-arms `ready`)
};

struct PeerTable {
    Mailbox* box[MAX_PEERS];  // box[rank] = the local mailbox
};

// by-value kernel argument; world == 0: no exchange
struct RangeExchange {
    PeerTable peers;
    int rank, world;
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One warp (all 32 lanes call it): mm[0] / mm[1] hold the ordered-int encodings of the LOCAL min / max on entry and of the
// min / max over all ranks on return.  Lane l talks to rank l: it stores (epoch, min) and (epoch, max) as two
// self-validating 64-bit words into this rank's slot of rank l's mailbox and spins on slot l of its own mailbox.
__device__ __forceinline__ void peer_minmax_warp(int* mm, const PeerTable& peers, int rank, int world) {
    const int lane = threadIdx.x & 31;
    Mailbox* mine = peers.box[rank];
    const unsigned e = mine->epoch + 1;
    const unsigned long long tag = (unsigned long long)e << 32;
    int lo = 0x7fffffff, hi = (int)0x80000000;
    if (lane < world) {
        const unsigned long long w_lo = tag | (unsigned)mm[0], w_hi = tag | (unsigned)mm[1];
        unsigned long long* dst = peers.box[lane]->slot[e & 1][rank];
        st_relaxed_sys_u64(dst, w_lo);
        st_relaxed_sys_u64(dst + 1, w_hi);
        const unsigned long long* src = mine->slot[e & 1][lane];
        const long long t0 = clock64();
        unsigned long long a, b;
        bool ok = true;
        const bool broken = *reinterpret_cast<volatile unsigned*>(&mine->timeouts) != 0;  // a peer was lost before: never wait again
        do {
            a = ld_relaxed_sys_u64(src);
            b = ld_relaxed_sys_u64(src + 1);
            if ((a >> 32) == e && (b >> 32) == e) break;
            if (broken || clock64() - t0 > 60000000000LL) {  // ~30 s: a peer never arrived -- do not hang the GPU for ever, flag the error
                ok = false;
                break;
            }
        } while (true);
        if (ok) {
            lo = (int)(unsigned)a;
            hi = (int)(unsigned)b;
        } else {
            atomicAdd(&mine->timeouts, 1u);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) {
        mm[0] = lo;
        mm[1] = hi;
        mine->epoch = e;
    }
}

// Prologue of a multi-block kernel that consumes the range: block 0's first warp runs the exchange and publishes the
// result, every other block waits for it (block 0 is always resident: it belongs to the first wave).  Call
// range_exchange_done() once per block after its last use of `mm`.
__device__ __forceinline__ void range_exchange_begin(int* mm, const RangeExchange& x) {
    if (x.world <= 1) return;
    Mailbox* mine = x.peers.box[x.rank];
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) {
            peer_minmax_warp(mm, x.peers, x.rank, x.world);
            __syncwarp();
            if (threadIdx.x == 0) {
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&mine->ready), "r"(1u) : "memory");
            }
        }
    } else if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->ready) : "memory");
        } while (v == 0);
    }
    __syncthreads();
}
__device__ __forceinline__ void range_exchange_done(const RangeExchange& x) {
    if (x.world <= 1) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        Mailbox* mine = x.peers.box[x.rank];
        if (atomicAdd(&mine->finished, 1u) == gridDim.x - 1) {  // last block of the launch: re-arm for the next one
            mine->finished = 0;
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&mine->ready), "r"(0u) : "memory");
        }
    }
}

}  // namespace rtb
