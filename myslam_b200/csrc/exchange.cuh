// Peer-memory exchange of the ray-sharded multi-GPU mapping (the reference is single-GPU,
// src/Mapper.py:348-350 is the optimiser step this replaces on several GPUs).
//
// Every rank keeps, at the SAME offsets of one symmetric allocation (NVLink peer-mapped, one per process):
//   parameter arena | gradient staging [world][slice] | published aux blocks (pose gradients, loss sums, counters)
//   | flag block
// The gradient arena itself stays in ORDINARY device memory: red.global.add into peer-mapped memory costs the
// fused backward +30 us per launch on B200 (tools/exchange_times.py), peer-mapped parameters cost nothing.
// After the local backward has reduced a rank's rays into its gradient arena, two back-to-back kernels per rank do
// reduce-scatter + Adam + all-gather + zero_grad over peer memory:
//   k_grad_push:     slice q of the local gradient arena is stored into rank q's staging row [rank] (P2P stores)
//                    and zeroed locally, for every q != rank
//   k_adam_exchange: barrier -> rank r sums its own slice and the staged rows in rank order (local loads) ->
//                    torch.optim.Adam on slice r (moments exist for the own slice only) -> the new parameters are
//                    stored into every rank's arena (P2P stores or one multimem.st through the NVSwitch) -> barrier
// No NCCL call and no separate Adam pass; replicas stay bit-identical because every parameter has one writer.
//
// Synchronisation: one LEADER CTA per rank shakes hands with the peers (world_size flag stores + polls, one
// system-scope fence); the other CTAs wait on / report to the leader through two local words with gpu-scope
// acquire/release, so a launch costs two system fences on the critical path instead of two per CTA (592 CTAs x
// MEMBAR.SYS serialise to ~18 us per barrier on B200, tools/exchange_times.py).  Only the leader ever waits at the
// end of the kernel, every other CTA exits when its slice is done, so the grid does not have to be co-resident.
// The staging rows need no second barrier: a rank only starts pushing call k+1 after the closing barrier of call
// k, which every rank reaches after it has consumed its staging.  The published counter / aux blocks alternate
// between two copies (the counter exchange has no closing barrier).  Flags carry a monotonically increasing epoch (never reset).  A poll
// that exceeds SPIN_LIMIT_NS raises the local status word and falls through, so a missing peer cannot hang the GPU.
#pragma once
#include "optim.cuh"

namespace eslam {

constexpr int MAX_PEERS = ESLAM_MAX_PEERS;
constexpr unsigned long long SPIN_LIMIT_NS = 4000000000ull;

struct PeerSync {
  unsigned* flags[MAX_PEERS];  // flag block of every rank: [slot][MAX_PEERS]
  int rank, world;
  unsigned epoch;
  int* status;                 // local, != 0 after a timeout
  unsigned long long* local;   // local words: [0] "go" epoch of the leader, [1] count of finished CTAs (monotonic)
  unsigned long long done_target;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_v4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Handshake of one CTA with the same CTA of every peer; all its threads call this.  What the CTA (and everything
// it has observed) wrote before is visible to the peers that pass the same barrier, and vice versa.
__device__ __forceinline__ void peer_barrier(const PeerSync& ps, int slot) {
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < ps.world) {
    st_release_sys(ps.flags[threadIdx.x] + slot * MAX_PEERS + ps.rank, ps.epoch);
    const unsigned* mine = ps.flags[ps.rank] + slot * MAX_PEERS + threadIdx.x;
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(mine) - ps.epoch) < 0) {
      if (global_ns() - t0 > SPIN_LIMIT_NS) {
        atomicExch(ps.status, 1);
        break;
      }
    }
  }
  __syncthreads();
}

// ---- loss normalisers: norm[i] = sum over ranks of counters[i]  (one CTA) ------------------------------------
struct CounterExchArgs {
  PeerSync ps;
  const int* local;          // this rank's counters
  int* pub[MAX_PEERS];       // every rank's published copy (this call's parity)
  int* norm;                 // local output
  int n;
};

__global__ void __launch_bounds__(32) k_exchange_counters(const __grid_constant__ CounterExchArgs a) {
  if ((int)threadIdx.x < a.n) a.pub[a.ps.rank][threadIdx.x] = a.local[threadIdx.x];
  peer_barrier(a.ps, 0);
  if ((int)threadIdx.x < a.n) {
    int acc = 0;
    for (int p = 0; p < a.ps.world; ++p) {
      int v;
      asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(a.pub[p] + threadIdx.x) : "memory");
      acc += v;
    }
    a.norm[threadIdx.x] = acc;
  }
}

// ---- reduce-scatter + Adam + all-gather ----------------------------------------------------------------------
struct AdamExchArgs {
  PeerSync ps;
  float4* p[MAX_PEERS];      // parameter arena of every rank
  float4* stage[MAX_PEERS];  // gradient staging of every rank: [world][slice_max] float4
  float4* g;                 // local gradient arena (ordinary memory)
  float4* mc_p;              // multicast alias of the parameter arenas (MULTIMEM only)
  float4 *m, *v;             // local moments (only the own slice is ever touched)
  AdamArgs adam;             // scalars + lr segments (p/g/m/v members unused)
  float* aux_local;             // this rank's pose-gradient block: published, then zeroed
  float* aux_pub[MAX_PEERS];    // every rank's published copy (this call's parity)
  float* aux_sum;               // local [n_aux]: sum over ranks
  int n_aux;
  double* auxd_local;           // this rank's loss sums
  double* auxd_pub[MAX_PEERS];
  double* auxd_sum;
  int n_auxd;
  int dbg;  // profiling aid (eslam_set_debug): bit3 skip the data loops, bit5 skip the barriers (results are
            // then meaningless)
};

constexpr int EXCH_THREADS = 256;
constexpr int SLOT_COUNTERS = 0, SLOT_IN = 1, SLOT_OUT = 2, N_SLOTS = 4;

__device__ __forceinline__ void slice_of(long long n4, int world, int r, long long& lo, long long& hi) {
  const long long base = n4 / world, rem = n4 % world;
  lo = r * base + (r < rem ? r : rem);
  hi = lo + base + (r < rem ? 1 : 0);
}

__device__ __forceinline__ void wait_local(const PeerSync& ps, int word, unsigned long long target) {
  if (threadIdx.x == 0) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_gpu(ps.local + word) < target) {
      if (global_ns() - t0 > SPIN_LIMIT_NS) {
        atomicExch(ps.status, 1);
        break;
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ long long slice_max(long long n4, int world) { return (n4 + world - 1) / world; }

// ---- reduce-scatter, push half: my contribution to every other rank's slice goes into its staging row [rank] --
__global__ void __launch_bounds__(EXCH_THREADS) k_grad_push(const __grid_constant__ AdamExchArgs a) {
  const int rank = a.ps.rank, world = a.ps.world;
  const long long n4 = a.adam.n >> 2;
  const long long smax = slice_max(n4, world);
  const long long stride = (long long)gridDim.x * EXCH_THREADS;
  const long long first = (long long)blockIdx.x * EXCH_THREADS + threadIdx.x;
  if (a.dbg & 8) return;
  const float4 z = f4_zero();
  for (int dq = 1; dq < world; ++dq) {
    const int q = (rank + dq) % world;  // start with a different peer on every rank: spreads the link load
    long long lo, hi;
    slice_of(n4, world, q, lo, hi);
    float4* dst = a.stage[q] + (long long)rank * smax - lo;
    // Whole, globally aligned groups of 32 float4 (128 parameters) per warp and trip.  A group whose gradient is all
    // zero on this rank is not sent: the owner keeps its staging rows zero (it clears what it consumes), and a
    // mapping window only ever touches part of the planes, so most of the NVLink volume disappears.
    const long long lo_a = lo & ~31ll, hi_r = (hi + 31) & ~31ll;
    for (long long i = lo_a + first; i < hi_r; i += stride) {
      const bool in = i >= lo && i < hi;
      const float4 v = in ? a.g[i] : z;
      const bool nz = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
      if (__ballot_sync(0xffffffffu, nz) == 0u) continue;
      if (in) {
        dst[i] = v;
        a.g[i] = z;
      }
    }
  }
}

// ---- reduce-scatter (sum half) + Adam + all-gather ---------------------------------------------------------------
// WMAX: compile-time bound of the world size (2, 4 or 8).
template <bool MULTIMEM, int WMAX>
__global__ void __launch_bounds__(EXCH_THREADS) k_adam_exchange(const __grid_constant__ AdamExchArgs a) {
  const int rank = a.ps.rank, world = a.ps.world;
  const int cta = blockIdx.x, nctas = gridDim.x;
  const long long n4 = a.adam.n >> 2;
  const long long smax = slice_max(n4, world);
  const long long stride = (long long)nctas * EXCH_THREADS;
  const long long first = (long long)cta * EXCH_THREADS + threadIdx.x;
  const bool leader = cta == 0;
  // ---- every rank's k_grad_push has completed (stream order on its side, handshake across ranks)
  if (leader) {
    // publish this rank's small blocks (pose gradients, loss sums) and clear the local ones for the next iteration
    for (int t = threadIdx.x; t < a.n_aux; t += EXCH_THREADS) {
      a.aux_pub[rank][t] = a.aux_local[t];
      a.aux_local[t] = 0.f;
    }
    for (int t = threadIdx.x; t < a.n_auxd; t += EXCH_THREADS) {
      a.auxd_pub[rank][t] = a.auxd_local[t];
      a.auxd_local[t] = 0.0;
    }
    if (!(a.dbg & 32)) peer_barrier(a.ps, SLOT_IN);
    if (threadIdx.x == 0) st_release_gpu(a.ps.local + 0, (unsigned long long)a.ps.epoch);
  } else if (!(a.dbg & 32)) {
    wait_local(a.ps, 0, (unsigned long long)a.ps.epoch);
  }
  long long lo, hi;
  slice_of(n4, world, rank, lo, hi);
  if (a.dbg & 8) hi = lo;
  float4* const p_own = a.p[rank];
  float4* const st_own = a.stage[rank] - lo;
  const float4 z = f4_zero();
  unsigned char* const touched = a.adam.touched;
  // Whole, globally aligned groups of 128 parameters per warp and trip (see k_grad_push).  A group in which no rank
  // has had a non-zero gradient since the optimiser was created still has m = v = 0: torch's update leaves p as it
  // is on every replica, so nothing is read beyond the gradients and nothing is stored or broadcast (k_adam).
  const long long lo_a = lo & ~31ll, hi_r = (hi + 31) & ~31ll;
  for (long long i = lo_a + first; i < hi_r; i += stride) {
    const bool in = i >= lo && i < hi;
    float4 part[WMAX];
    bool nz = false;
#pragma unroll
    for (int q = 0; q < WMAX; ++q)
      if (q < world) {
        part[q] = !in ? z : (q == rank) ? a.g[i] : ld_sys_v4(st_own + q * smax + i);
        nz = nz || part[q].x != 0.f || part[q].y != 0.f || part[q].z != 0.f || part[q].w != 0.f;
      }
    const bool any = __ballot_sync(0xffffffffu, nz) != 0u;
    const bool was = touched ? touched[i >> 5] != 0 : true;
    if (!any && !was) continue;
    if (touched && !was && (threadIdx.x & 31) == 0) touched[i >> 5] = 1;
    if (!in) continue;
    if (any) {  // consumed: the gradient arena and the staging rows go back to zero
      a.g[i] = z;
#pragma unroll
      for (int q = 0; q < WMAX; ++q)
        if (q < world && q != rank) st_own[q * smax + i] = z;
    }
    float4 p = p_own[i], m = a.m[i], v = a.v[i];
    float4 g = part[0];
#pragma unroll
    for (int q = 1; q < WMAX; ++q)
      if (q < world) g = f4_add(g, part[q]);  // fixed rank order: the sum does not depend on who owns the slice
    const long long e = i << 2;
    int seg = 0;
    while (seg < a.adam.n_seg - 1 && e >= a.adam.seg_end[seg]) ++seg;
    const float ss = a.adam.seg_step[seg];
    adam_one(p.x, g.x, m.x, v.x, a.adam, ss);
    adam_one(p.y, g.y, m.y, v.y, a.adam, ss);
    adam_one(p.z, g.z, m.z, v.z, a.adam, ss);
    adam_one(p.w, g.w, m.w, v.w, a.adam, ss);
    a.m[i] = m;
    a.v[i] = v;
    if (MULTIMEM) {
      multimem_st_v4(a.mc_p + i, p);
    } else {
#pragma unroll
      for (int q = 0; q < WMAX; ++q)
        if (q < world) a.p[q][i] = p;
    }
  }
  // ---- small replicated sums (pose gradients, loss terms): every rank adds all ranks' blocks in rank order
  if (leader) {
    for (int t = threadIdx.x; t < a.n_aux; t += EXCH_THREADS) {
      float acc = 0.f;
      for (int p = 0; p < world; ++p) {
        float x;
        asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x) : "l"(a.aux_pub[p] + t) : "memory");
        acc += x;
      }
      a.aux_sum[t] = acc;
    }
    for (int t = threadIdx.x; t < a.n_auxd; t += EXCH_THREADS) {
      double acc = 0.0;
      for (int p = 0; p < world; ++p) {
        double x;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(a.auxd_pub[p] + t) : "memory");
        acc += x;
      }
      a.auxd_sum[t] = acc;
    }
  }
  if (a.dbg & 32) {  // profiling: keep the finished-CTA count consistent, skip the waits
    if (threadIdx.x == 0) atomicAdd(a.ps.local + 1, 1ull);
    return;
  }
  // ---- this CTA's parameter stores are performed system-wide; report to the leader and leave
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(a.ps.local + 1, 1ull);
  }
  if (!leader) return;
  // ---- leader: all local CTAs have stored their slice into every rank -> tell the peers, wait for theirs
  wait_local(a.ps, 1, a.ps.done_target);
  peer_barrier(a.ps, SLOT_OUT);
}

}  // namespace eslam
