// Peer-memory exchange of the ray-sharded multi-GPU mapping (the reference is single-GPU,
// src/Mapper.py:348-350 is the optimiser step this replaces on several GPUs).
//
// Every rank keeps, at the SAME offsets of one symmetric allocation (NVLink peer-mapped, one per process):
//   parameter arena | gradient-image staging [world][slice] | published aux blocks (decoder gradients, pose
//   gradients, loss sums, counters) | flag block
// The gradient images themselves stay in ORDINARY device memory: red.global.add into peer-mapped memory costs the
// fused backward +30 us per launch on B200 (tools/exchange_times.py), peer-mapped parameters cost nothing.
// No NCCL call and no separate Adam pass; replicas stay bit-identical because every parameter has one writer
// (planes: the owner of the tile; decoders: every rank computes the same sum in the same order).
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
#include "qplane.cuh"

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
    int v[MAX_PEERS];
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < a.ps.world) asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v[p]) : "l"(a.pub[p] + threadIdx.x) : "memory");
    int acc = 0;
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < a.ps.world) acc += v[p];
    a.norm[threadIdx.x] = acc;
  }
}

// ---- small replicated sums on their own (pose gradients, loss terms): publish, handshake, add in rank order, clear the
//      local block.  One CTA; used on a side stream so that the pose step and the next iteration's ray sampling need
//      not wait for the plane exchange.
struct AuxExchArgs {
  PeerSync ps;
  float* aux_local;
  float* aux_pub[MAX_PEERS];
  float* aux_sum;
  int n_aux;
  double* auxd_local;
  double* auxd_pub[MAX_PEERS];
  double* auxd_sum;
  int n_auxd;
};

constexpr int SLOT_AUX = 3;

__global__ void __launch_bounds__(256) k_exchange_aux(const __grid_constant__ AuxExchArgs a) {
  const int rank = a.ps.rank, world = a.ps.world;
  for (int t = threadIdx.x; t < a.n_aux; t += 256) {
    a.aux_pub[rank][t] = a.aux_local[t];
    a.aux_local[t] = 0.f;
  }
  for (int t = threadIdx.x; t < a.n_auxd; t += 256) {
    a.auxd_pub[rank][t] = a.auxd_local[t];
    a.auxd_local[t] = 0.0;
  }
  peer_barrier(a.ps, SLOT_AUX);
  for (int t = threadIdx.x; t < a.n_aux; t += 256) {
    float x[MAX_PEERS];
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x[p]) : "l"(a.aux_pub[p] + t) : "memory");
    float acc = 0.f;
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) acc += x[p];
    a.aux_sum[t] = acc;
  }
  for (int t = threadIdx.x; t < a.n_auxd; t += 256) {
    double x[MAX_PEERS];
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x[p]) : "l"(a.auxd_pub[p] + t) : "memory");
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < MAX_PEERS; ++p)
      if (p < world) acc += x[p];
    a.auxd_sum[t] = acc;
  }
}

// ---- reduce-scatter of the gradient images + plane Adam + all-gather, then the decoders' replicated step ----------
// The mapping backward (qbwd.cuh) leaves a rank's plane gradients as 16-channel gradient images (13.5 MB for room0:
// half of a parameter-form gradient arena) and its decoder gradients in the gradient arena's decoder block.
// Ownership is by TILES of the optimiser tail (qplane.cuh: QA_TILE texels): rank r owns a contiguous range of
// tiles, i.e. a contiguous slice of the images.
//   k_gq_push          every rank stores the peers' slices of its gradient image into their staging rows [rank]
//                      (P2P stores) and zeroes them locally; all-zero groups of 128 floats are not sent
//   k_q_adam_exchange  barrier -> the tail (q_adam_tile) on the owned tiles, whose gradient rows are the sum of the
//                      local image and the staged rows in rank order and whose updated texels go to EVERY rank's
//                      arena -> the leader publishes this rank's decoder gradients (dW1 of the owned tiles, the
//                      backward's other decoder gradients) -> barrier
//   k_dec_adam_peers   every rank sums all ranks' published decoder gradients in rank order and takes the
//                      decoders' Adam step (replicated: identical on every rank)
struct QExchArgs {
  PeerSync ps;
  QAdamArgs q;
  long long lo4[MAX_PEERS + 1];  // first float4 of every rank's slice of the gradient images
  float4* stage[MAX_PEERS];      // staging of every rank: [world][smax] float4
  float* dec_pub[MAX_PEERS];     // every rank's published decoder gradients (this call's parity), DEC_N floats
  float* aux_local;              // this rank's pose-gradient block: published, then zeroed
  float* aux_pub[MAX_PEERS];     // every rank's published copy (this call's parity)
  float* aux_sum;                // local [n_aux]: sum over ranks
  int n_aux;
  double* auxd_local;            // this rank's loss sums
  double* auxd_pub[MAX_PEERS];
  double* auxd_sum;
  int n_auxd;
  int dbg;  // profiling aid (eslam_set_debug): bit3 skip the data loops, bit5 skip the barriers (results are
            // then meaningless)
};

constexpr int EXCH_THREADS = 256;
constexpr int SLOT_COUNTERS = 0, SLOT_IN = 1, SLOT_OUT = 2, N_SLOTS = 4;

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

// ---- reduce-scatter, push half: my contribution to every other rank's slice goes into its staging row [rank] --
__global__ void __launch_bounds__(EXCH_THREADS) k_gq_push(const __grid_constant__ QExchArgs a) {
  const int rank = a.ps.rank, world = a.ps.world;
  const long long stride = (long long)gridDim.x * EXCH_THREADS;
  const long long first = (long long)blockIdx.x * EXCH_THREADS + threadIdx.x;
  if (a.dbg & 8) return;
  const float4 z = f4_zero();
  for (int dq = 1; dq < world; ++dq) {
    const int q = (rank + dq) % world;  // start with a different peer on every rank: spreads the link load
    const long long lo = a.lo4[q], hi = a.lo4[q + 1];
    float4* dst = a.stage[q] + (long long)rank * a.q.smax - lo;
    // Whole, globally aligned groups of 32 float4 (8 texels) per warp and trip.  A group whose gradient is all
    // zero on this rank is not sent: the owner keeps its staging rows zero (it clears what it consumes), and a
    // mapping window only ever touches part of the planes, so most of the NVLink volume disappears.
    const long long lo_a = lo & ~31ll, hi_r = (hi + 31) & ~31ll;
    for (long long i = lo_a + first; i < hi_r; i += stride) {
      const bool in = i >= lo && i < hi;
      const float4 v = in ? a.q.gq4[i] : z;
      const bool nz = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
      if (__ballot_sync(0xffffffffu, nz) == 0u) continue;
      if (in) {
        dst[i] = v;
        a.q.gq4[i] = z;
      }
    }
  }
}

// WMAX: compile-time bound of the world size (2, 4 or 8).  Grid = the tiles this rank owns.
template <int WMAX>
__global__ void __launch_bounds__(QA_THREADS, 3) k_q_adam_exchange(const __grid_constant__ QExchArgs a) {
  __shared__ SmemQAdam sm;
  const int rank = a.ps.rank, world = a.ps.world;
  const bool leader = blockIdx.x == 0;
  // ---- every rank's k_gq_push has completed (stream order on its side, handshake across ranks)
  if (leader) {
    // publish this rank's small blocks (pose gradients, loss sums) and clear the local ones for the next iteration
    for (int t = threadIdx.x; t < a.n_aux; t += QA_THREADS) {
      a.aux_pub[rank][t] = a.aux_local[t];
      a.aux_local[t] = 0.f;
    }
    for (int t = threadIdx.x; t < a.n_auxd; t += QA_THREADS) {
      a.auxd_pub[rank][t] = a.auxd_local[t];
      a.auxd_local[t] = 0.0;
    }
    if (!(a.dbg & 32)) peer_barrier(a.ps, SLOT_IN);
    if (threadIdx.x == 0) st_release_gpu(a.ps.local + 0, (unsigned long long)a.ps.epoch);
  } else if (!(a.dbg & 32)) {
    wait_local(a.ps, 0, (unsigned long long)a.ps.epoch);
  }
  if (!(a.dbg & 8)) q_adam_tile<WMAX>(a.q, a.q.unit_lo + blockIdx.x, sm);
  // ---- small replicated sums (pose gradients, loss terms): every rank adds all ranks' blocks in rank order (the peer
  //      loads of an element are independent: one NVLink round trip, not world_size of them)
  if (leader) {
    for (int t = threadIdx.x; t < a.n_aux; t += QA_THREADS) {
      float x[MAX_PEERS];
#pragma unroll
      for (int p = 0; p < MAX_PEERS; ++p)
        if (p < world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x[p]) : "l"(a.aux_pub[p] + t) : "memory");
      float acc = 0.f;
#pragma unroll
      for (int p = 0; p < MAX_PEERS; ++p)
        if (p < world) acc += x[p];
      a.aux_sum[t] = acc;
    }
    for (int t = threadIdx.x; t < a.n_auxd; t += QA_THREADS) {
      double x[MAX_PEERS];
#pragma unroll
      for (int p = 0; p < MAX_PEERS; ++p)
        if (p < world) asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x[p]) : "l"(a.auxd_pub[p] + t) : "memory");
      double acc = 0.0;
#pragma unroll
      for (int p = 0; p < MAX_PEERS; ++p)
        if (p < world) acc += x[p];
      a.auxd_sum[t] = acc;
    }
  }
  if (a.dbg & 32) {  // profiling: keep the finished-CTA count consistent, skip the waits
    if (threadIdx.x == 0) atomicAdd(a.ps.local + 1, 1ull);
    return;
  }
  // ---- this CTA's parameter stores and dW1 reductions are performed system-wide; report to the leader and leave
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    atomicAdd(a.ps.local + 1, 1ull);
  }
  if (!leader) return;
  // ---- leader: all local CTAs are done -> publish the decoder gradients, tell the peers, wait for theirs
  wait_local(a.ps, 1, a.ps.done_target);
  for (int t = threadIdx.x; t < DEC_N; t += QA_THREADS) {
    a.dec_pub[rank][t] = __ldcg(a.q.gdec + t);
    a.q.gdec[t] = 0.f;
  }
  peer_barrier(a.ps, SLOT_OUT);
}

// the decoders' Adam step on the sum of all ranks' published gradients (torch/optim/adam.py, exact operation order)
struct DecPeersArgs {
  int world;
  const float* dec_pub[MAX_PEERS];
  float *p, *m, *v;  // local decoder block of the parameter / moment arenas
  AdamArgs adam;     // seg_step[0] = lr / (1 - beta1^t)
};

// one element per thread: the peer loads of a thread are independent and all in flight at once, so the kernel costs
// one NVLink round trip (a single CTA striding over the 2 700 floats paid eleven of them back to back: 33 us)
__global__ void __launch_bounds__(256) k_dec_adam_peers(const __grid_constant__ DecPeersArgs a) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= DEC_N) return;
  float x[MAX_PEERS];
#pragma unroll
  for (int q = 0; q < MAX_PEERS; ++q)
    if (q < a.world) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(x[q]) : "l"(a.dec_pub[q] + t) : "memory");
  float g = x[0];
#pragma unroll
  for (int q = 1; q < MAX_PEERS; ++q)
    if (q < a.world) g += x[q];  // rank order: identical on every rank
  float p = a.p[t], m = a.m[t], v = a.v[t];
  adam_one(p, g, m, v, a.adam, a.adam.seg_step[0]);
  a.p[t] = p;
  a.m[t] = m;
  a.v[t] = v;
}

}  // namespace eslam
