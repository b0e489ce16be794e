// spconv_tc.cu -- sparse convolution on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One CTA owns a tile of 128 output rows x NT <= 128 output columns (and, with split-K, a contiguous share of the
// (kernel offset, 32-channel slice) iterations). For every kernel offset with at least one active pair in the tile
// and every 32-channel slice of Cin:
//   * the producer threads gather one input row slice each. Inputs are PACKED rows: activated (the consumer's
//     BatchNorm(eval)+ReLU) and split once by the PRODUCING convolution's epilogue or by sgb_act_pack. A value x is carried
//     as two fp16 numbers hi = fp16(x), lo = fp16((x - hi) * 2^kLoShift); both go straight into TENSOR MEMORY with
//     tcgen05.st (A operand from TMEM: lane = row, 16 columns of hi pairs + 16 columns of lo pairs per stage) -- the
//     shared-memory pipe only carries the weights;
//   * the pre-packed weight slice (same split, core-matrix order, done once on the host side) is streamed by one
//     elected lane with TMA bulk copies into a ring of stages; the weights of the whole convolution are pulled
//     into L2 with cp.async.bulk.prefetch at kernel start because every CTA walks them in the same order (so they
//     would otherwise always be cold);
//   * one thread issues the error-compensated products as tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16)
//         D[:, 0:2nt]  += A_hi * [B_hi | B_lo]      D[:, nt:2nt] += A_lo * B_hi
//     accumulating fp32 in TMEM, then tcgen05.commit's the stages back to their producers through mbarriers.
// The epilogue reads the accumulator with tcgen05.ld, sums the two column blocks, adds bias / residual and writes
// strided fp32 rows (the U-Net concat buffer) and/or the PACKED rows its consumer reads (that consumer's BatchNorm+ReLU
// folded in: an intermediate with a single consumer never exists in fp32). With split-K (deep U-Net levels: a handful of row tiles, megabytes of
// weights) the CTAs of one thread-block cluster hold partial tiles; ranks > 0 park theirs in their own shared memory
// and rank 0 sums them in rank order over distributed shared memory -- deterministic, no workspace, no atomics.
// Dropping lo*lo leaves a relative error of ~2^-22 per product, i.e. fp32-grade (the north star's 1e-4 over ~40
// sequential convolutions rules out plain TF32/fp16).
#include <algorithm>
#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace sgb {

// spconv_ss.cu: persistent shared-memory-ring kernel (large levels)
bool spconv_ss_plan(int K, int Mout, int Cin, int Cout, int sms, int *out);
int spconv_ss_launch(const float *d_in_pk, int in_stride, const int32_t *d_map, int K, int Mout, const float *d_Wp, int Cin,
                     int Cout, const float *d_residual, int res_stride, int res_off, const float *d_bias, float *d_out,
                     int out_stride, int out_off, float *d_pk_out, int pk_stride, int pk_coff, const float *d_pk_scale,
                     const float *d_pk_shift, int pk_relu, int pk_fill, int *d_oflow, int sms, bool *attr_set, cudaStream_t stream);

constexpr int TC_ROWS = 128;
constexpr int TC_KC = 32;  // channels per stage
constexpr int TC_THREADS = 320;  // 8 producer warps + 1 MMA warp + 1 weight-loader warp
// The remainder lo = x - fp16(x) is carried as fp16(lo * 2^kLoShift); the correction columns are scaled back by
// 2^-kLoShift when they are added to the main product. With shift 0 (round 1) lo was a subnormal fp16 for |x| < ~0.1
// (2^-25 absolute error) and the per-element 1e-4 bar was missed at C = 224 (2.9e-4, GPU call 3 of round 2); 11 keeps
// lo normal down to |x| = 2^-14 (sgb_spconv_lo_shift() tells the host weight packer).
constexpr int kLoShift = 11;
constexpr float kLoScale = (float)(1 << kLoShift), kLoInv = 1.0f / kLoScale;
static_assert(kLoShift == 11, "spconv_ss.cu hard-codes the remainder scale 2^11");
constexpr int BAR_FULL = 0, BAR_FREE = 3, BAR_BFULL = 9;  // per pair stage (<= 3); bars[8] = accumulator done

struct TcArgs {
  const float *in; int in_stride, in_off;  // packed rows (words)
  const int32_t *map; int K, Mout;
  const float *Wp;  // packed fp16 [K][nkc][4 chunks][2 (hi,lo)][N][8 halves]
  int Cin, N, Cout;        // N = Cout rounded up to 16
  int NT;                  // columns per CTA (multiple of 16); gridDim.y = ceil(N / NT)
  const float *residual; int res_stride, res_off;
  const float *bias;
  float *out; int out_stride, out_off;   // optional fp32 rows
  uint32_t *pk; int pk_stride, pk_coff;  // optional packed rows: row stride in words, first channel (multiple of 8)
  const float *pk_scale, *pk_shift; int pk_relu;  // per OUTPUT channel (nullptr: identity): the consumer's BatchNorm+ReLU
  int pk_fill;                           // zero the rest of a half-written last 32-channel chunk
  int *oflow;                            // device flag: a packed value beyond the fp16 range
  int nstages, nbstages, tmem_cols, tmem_acols;  // A ring depth (TMEM), weight ring depth (smem), TMEM columns, first A column
  int ksplit;      // CTAs per cluster along z sharing one tile's iterations (1 = no split)
};

// Warp roles: warps 0-3 = producer group 0, warps 4-7 = producer group 1 (one output row per thread; group g fills slot g
// of every iteration pair), warp 8 = MMA issuer (warp-uniform, one elected lane), warp 9 = weight loader (TMA bulk copies).
// Barriers per pair stage: full (8 producer-warp arrivals), bfull (weights: expect_tx), free (tcgen05.commit);
// bars[8]: accumulator complete.
__global__ void __launch_bounds__(TC_THREADS, 2) spconv_tc_kernel(TcArgs p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[12];  // [0..2] A full, [3..5] pair free, [8] done, [9..11] weights full
  __shared__ uint32_t s_tmem;
  __shared__ unsigned int s_mask;
  __shared__ int s_list[32];
  __shared__ int s_nact;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);  // provably warp-uniform copy for the role dispatch
  const bool producer = warp < 8;
  const int grp = warp >> 2;          // producer group (0/1)
  const int r = tid & (TC_ROWS - 1);  // row of this producer thread
  const int row0 = blockIdx.x * TC_ROWS;
  const int my_row = row0 + r;
  const bool row_ok = producer && my_row < p.Mout;
  const int N = p.N, NT = p.NT;
  const int n0 = blockIdx.y * NT;
  const int nt = min(NT, N - n0);                     // columns of this CTA (multiple of 16)
  const uint32_t b_bytes = (uint32_t)NT * TC_KC * 2;  // NT * 64 B (fp16)
  const uint32_t bstage_bytes = 2 * b_bytes;    // weight ring stage: hi + lo
  const int NS = p.nstages;   // A stages in TMEM (32 columns each) == weight stages in shared memory; even
  unsigned char *bring = smem;
  int32_t *map_s = reinterpret_cast<int32_t *>(bring + (size_t)NS * bstage_bytes);  // [K][128] (only when p.map)

  if (tid == 0) {
    for (int i = 0; i < NS / 2; i++) {
      mbar_init(smem_u32(&bars[BAR_FULL + i]), 8);  // one arrival per producer warp
      mbar_init(smem_u32(&bars[BAR_FREE + i]), 1);
      mbar_init(smem_u32(&bars[BAR_BFULL + i]), 1);
    }
    mbar_init(smem_u32(&bars[8]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mask = 0u;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  // ---- rulebook slice of this tile -> shared memory; which kernel offsets have any active pair ------------
  if (p.map) {
    unsigned int flags = 0u;
    if (producer) {
      // all of this thread's rulebook entries are requested before the first one is used (one memory latency, not K/2)
      for (int ob = grp; ob < p.K; ob += 28) {
        int srcv[14];
#pragma unroll
        for (int j = 0; j < 14; j++) {
          const int o = ob + 2 * j;
          srcv[j] = (o < p.K && row_ok) ? __ldg(&p.map[(size_t)o * p.Mout + my_row]) : -1;
        }
#pragma unroll
        for (int j = 0; j < 14; j++) {
          const int o = ob + 2 * j;
          if (o < p.K) {
            map_s[o * TC_ROWS + r] = srcv[j];
            if (srcv[j] >= 0) flags |= 1u << o;
          }
        }
      }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0 && flags) atomicOr(&s_mask, flags);
  }
  __syncthreads();
  if (tid < 32) {  // compact the active offsets in ascending order
    if (p.map) {
      const unsigned int m = s_mask;
      if (tid < p.K && (m >> tid & 1u)) s_list[__popc(m & ((1u << tid) - 1u))] = tid;
      if (tid == 0) s_nact = __popc(m);
    } else if (tid == 0) {
      s_list[0] = 0;
      s_nact = 1;
    }
  }
  __syncthreads();
  const int nkc = (p.Cin + TC_KC - 1) / TC_KC;
  // split-K: the cluster's CTAs take contiguous shares of the tile's (offset, slice) iterations
  const int S = p.ksplit, z = blockIdx.z;
  const int total_all = s_nact * nkc;
  const int i_beg = (int)((long long)total_all * z / S), i_end = (int)((long long)total_all * (z + 1) / S);
  const int total = i_end - i_beg;
  const int a0 = i_beg / nkc, k0 = i_beg - a0 * nkc;  // first iteration of this CTA

  // Pipeline unit = a PAIR of consecutive iterations (2P, 2P+1): producer group g fills A stage (ps, g), the loader
  // brings both weight slices under one transaction barrier, and the MMA thread pays its fixed costs (two barrier
  // waits, one commit) once per pair and issues up to 8 back-to-back MMAs, which is what lets them pipeline in the
  // tensor core. Per pair stage ps: full[ps] (8 producer-warp arrivals), bfull[ps] (weights landed),
  // free[ps] (tcgen05.commit: both the A stages and the weight stages of the pair can be overwritten).
  const int npairs = (total + 1) >> 1;
  const int NP = NS >> 1;
  if (producer) {
    // ---- packed input: the gather is pure data movement, so the registers hold TWO future iterations of this thread's
    //      row (4 iterations ahead of the MMA warp counting both groups) -- the L2 latency of the gather is covered
    //      without shared memory.
    uint32_t va[32], vb[32];
    int ia = a0, ikc = k0 + grp;  // issue cursor: (offset list position, channel slice) of the next own iteration to load
    while (ikc >= nkc) { ikc -= nkc; ia++; }
    int nload = grp;        // CTA-local iteration index of the next load
    // One gathered row slice = one 128-byte line, every lane a different line: the loads are tag-bound in L1, so the
    // slice is fetched with four 256-bit loads (LDG.E.256) instead of eight 128-bit ones.
    auto load = [&](uint32_t (&buf)[32]) {
      if (nload < total) {
        const int o = s_list[ia];
        const int src = p.map ? map_s[o * TC_ROWS + r] : (row_ok ? my_row : -1);
        if (src >= 0) {
          const float *rp = p.in + (size_t)src * p.in_stride + p.in_off + ikc * TC_KC;
#pragma unroll
          for (int q = 0; q < 4; q++)
            asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(buf[8 * q + 0]), "=r"(buf[8 * q + 1]), "=r"(buf[8 * q + 2]), "=r"(buf[8 * q + 3]),
                           "=r"(buf[8 * q + 4]), "=r"(buf[8 * q + 5]), "=r"(buf[8 * q + 6]), "=r"(buf[8 * q + 7])
                         : "l"(rp + 8 * q));
        } else {
#pragma unroll
          for (int q = 0; q < 32; q++) buf[q] = 0u;
        }
      }
      nload += 2;
      ikc += 2;
      while (ikc >= nkc) { ikc -= nkc; ia++; }
    };
    int ps = 0, u = 0, P = 0;
    auto consume = [&](uint32_t (&buf)[32]) {
      const bool work = 2 * P + grp < total;  // an odd tail leaves group 1 without a slice: it only arrives
      if (u >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((u - 1) & 1));
      if (work) {
        const uint32_t ta = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(p.tmem_acols + (2 * ps + grp) * 32);
        tmem_st32(ta, buf);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      // one arrival per WARP (every lane has completed and fenced its own store): 256 per-thread arrivals on one
      // mbarrier serialise in the shared-memory atomic unit and were the floor of the whole pipeline
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars[BAR_FULL + ps]));
      // Refill this buffer (own iteration +2) only AFTER the arrive: tcgen05.wait::st / tcgen05.fence compile to
      // FENCE.VIEW.ASYNC, which also waits for every outstanding global load of the thread -- issued before the
      // fence, the gather's L2 latency was paid in full on every iteration instead of overlapping the other buffer.
      if (work) load(buf);
      P++;
      if (++ps == NP) { ps = 0; u++; }
    };
    load(va);
    load(vb);
    for (int j = 0; j < npairs; j += 2) {
      consume(va);
      if (j + 1 < npairs) consume(vb);
    }
  } else if (warp == 9) {
    // ---- weight loader: one elected lane streams the packed slices with TMA bulk copies (cp.async.bulk, async proxy:
    //      no generic->async fence needed); completion is signalled on the pair's mbarrier by complete_tx. Global
    //      layout [chunk q][hi|lo][N][16 B] == shared layout [q][hi nt | lo nt][16 B] when the CTA owns all N columns,
    //      so a whole slice is ONE bulk copy; with a column split it is one copy per (q, part), one lane each.
    int ps = 0, ub = 0, kc = k0, a_idx = a0;
    for (int P = 0; P < npairs; P++) {
      if (ub >= 1) mbar_wait(smem_u32(&bars[BAR_FREE + ps]), (uint32_t)((ub - 1) & 1));
      const int nit = min(2, total - 2 * P);
      const uint32_t bar = smem_u32(&bars[BAR_BFULL + ps]);
      // slot 0 / slot 1 of the pair: offset, slice, 16-channel k-steps
      const int o0 = s_list[a_idx], kc0 = kc;
      int kc1 = kc + 1, a1 = a_idx;
      if (kc1 == nkc) { kc1 = 0; a1++; }
      const int o1 = (nit > 1) ? s_list[a1] : 0;
      const int ks0 = (min(TC_KC, p.Cin - kc0 * TC_KC) + 15) >> 4;
      const int ks1 = (nit > 1) ? (min(TC_KC, p.Cin - kc1 * TC_KC) + 15) >> 4 : 0;
      const int nseg0 = 4 * ks0, nseg1 = 4 * ks1;  // (chunk, part) segments of nt*16 bytes
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)(nseg0 + nseg1) * (uint32_t)nt * 16u;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
      }
      __syncwarp();
      const float4 *g0 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o0 * nkc + kc0) * (size_t)N * 8;
      const float4 *g1 = reinterpret_cast<const float4 *>(p.Wp) + ((size_t)o1 * nkc + kc1) * (size_t)N * 8;
      const uint32_t bb0 = smem_u32(bring + (size_t)(2 * ps) * bstage_bytes), bb1 = bb0 + bstage_bytes;
      if (nt == N) {
        if (lane == 0)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bb0), "l"(g0), "r"((uint32_t)nseg0 * (uint32_t)nt * 16u), "r"(bar) : "memory");
        if (lane == 1 && nit > 1)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(bb1), "l"(g1), "r"((uint32_t)nseg1 * (uint32_t)nt * 16u), "r"(bar) : "memory");
      } else if (lane < nseg0 + nseg1) {  // column split: one bulk copy per (chunk, part) segment
        const bool second = lane >= nseg0;
        const int sg = second ? lane - nseg0 : lane;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((second ? bb1 : bb0) + (uint32_t)(sg * nt) * 16u), "l"((second ? g1 : g0) + (size_t)sg * N + n0),
                     "r"((uint32_t)nt * 16u), "r"(bar)
                     : "memory");
      }
      if (++ps == NP) { ps = 0; ub++; }
      kc += 2;
      while (kc >= nkc) { kc -= nkc; a_idx++; }
    }
  } else if (warp_u == 8) {
    // ---- MMA issuer: per 16-channel k-step two instructions
    //        D[:, 0:2nt]  += A_hi * [B_hi | B_lo]     (N = 2nt)
    //        D[:, nt:2nt] += A_lo *  B_hi             (N = nt)
    //      so columns [0,nt) hold hi*hi and [nt,2nt) the two correction products (summed in the epilogue).
    //      The WHOLE warp runs this loop and one elected lane issues: with warp-uniform control flow and operands the
    //      compiler keeps descriptors in uniform registers and emits bare UTCHMMA; issued from a divergent
    //      `lane == 0` branch every instruction was wrapped in an R2UR + VOTEU/ELECT/BRA.U.ANY waterfall loop, and that
    //      single thread's instruction stream (not the tensor pipe) paced the whole kernel.
    //      Measured floor (scripts/umma_rate.py, round 2): 151 cycles per k-step (both instructions) for nt = 32..96,
    //      200 at nt = 128, independent of N below that -- 302 cycles per 32-channel iteration and SM is the tensor
    //      pipe's issue floor for this operand shape (M = 128, K = 16, A from TMEM).
    // instruction descriptor: D = f32 (1 << 4), A = B = f16 (format 0), both K-major, N >> 3, M >> 4
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const int total_u = __shfl_sync(0xffffffffu, total, 0);
    const int npairs_u = (total_u + 1) >> 1;
    const int k0_u = __shfl_sync(0xffffffffu, k0, 0);
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t idesc2 = (1u << 4) | ((uint32_t)((2 * nt) >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t idesc1 = (1u << 4) | ((uint32_t)(nt >> 3) << 17) | ((uint32_t)(TC_ROWS >> 4) << 24);
    const uint32_t b_lbo = (uint32_t)(2 * nt) * 16;
    const uint64_t b_step = (uint64_t)((2 * b_lbo) >> 4);
    const uint32_t bring_u = smem_u32(bring);
    const uint32_t bars_u = smem_u32(&bars[0]);
    const uint32_t acol0 = tmem_u + (uint32_t)p.tmem_acols;
    uint32_t first = 0u;  // 0 for the very first MMA (overwrite), then 1
    int ps = 0, kc = k0_u;
    uint32_t par = 0u;
    for (int P = 0; P < npairs_u; P++) {
      mbar_wait(bars_u + 8u * (uint32_t)(BAR_BFULL + ps), par);
      mbar_wait(bars_u + 8u * (uint32_t)(BAR_FULL + ps), par);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int nit = min(2, total_u - 2 * P);
      for (int h = 0; h < nit; h++) {
        const int kvalid = min(TC_KC, p.Cin - kc * TC_KC);
        const int ksteps = (kvalid + 15) >> 4;
        const uint32_t sbt = bring_u + (uint32_t)(2 * ps + h) * bstage_bytes;
        uint64_t dbb = umma_desc(sbt, b_lbo, 128);
        uint32_t ah = acol0 + (uint32_t)((2 * ps + h) * 32), al = ah + 16u;
        for (int ks = 0; ks < ksteps; ks++) {
          if (leader) {
            umma_f16_ts(tmem_u, ah, dbb, idesc2, first);
            umma_f16_ts(tmem_u + (uint32_t)nt, al, dbb, idesc1, 1u);
          }
          first = 1u;
          ah += 8u; al += 8u; dbb += b_step;
        }
        if (++kc == nkc) kc = 0;
      }
      if (leader) umma_commit(bars_u + 8u * (uint32_t)(BAR_FREE + ps));  // frees both A stages and both weight stages of the pair
      __syncwarp();
      if (++ps == NP) { ps = 0; par ^= 1u; }
    }
    if (leader && total_u > 0) umma_commit(bars_u + 8u * 8u);
    __syncwarp();
  }
  // ---- epilogue: TMEM lane = output row; warp w reads lanes 32*(w%4).., column half w/4 -----------------
  const bool split = S > 1;
  const uint32_t pbuf = smem_u32(bring);  // split-K partial tile [nt columns][128 rows] f32 (the weight ring is idle by now)
  if (split) {
    if (producer && z > 0) {
      if (total > 0) {
        mbar_wait(smem_u32(&bars[8]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const int lane_grp = warp & 3, chalf = warp >> 2;
      const int cbeg = chalf * (nt >> 1), cend = cbeg + (nt >> 1);
      for (int cb = cbeg; cb < cend; cb += 8) {
        uint32_t v[8], c[8];
        if (total > 0) {
          uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cb;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
                       : "r"(taddr + (uint32_t)nt));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = c[e] = 0u;
        }
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const float x = fmaf(__uint_as_float(c[e]), kLoInv, __uint_as_float(v[e]));
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(pbuf + (uint32_t)(((cb + e) * TC_ROWS + lane_grp * 32 + lane) * 4)), "f"(x) : "memory");
        }
      }
    }
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (producer && (!split || z == 0)) {
    if (total > 0) {
      mbar_wait(smem_u32(&bars[8]), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int lane_grp = warp & 3, chalf = warp >> 2;
    const int row = row0 + lane_grp * 32 + lane;
    const int cbeg = chalf * (nt >> 1), cend = cbeg + (nt >> 1);
    for (int cb = cbeg; cb < cend; cb += 8) {
      uint32_t v[8], c[8];
      if (total > 0) {
        uint32_t taddr = tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)cb;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
                     : "r"(taddr + (uint32_t)nt));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __float_as_uint(fmaf(__uint_as_float(c[e]), kLoInv, __uint_as_float(v[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = 0u;
      }
      if (split) {  // add the other ranks' partial tiles in rank order (distributed shared memory)
        for (int zz = 1; zz < S; zz++) {
          uint32_t rbase;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbase) : "r"(pbuf), "r"(zz));
          float t[8];
#pragma unroll
          for (int e = 0; e < 8; e++)
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(t[e]) : "r"(rbase + (uint32_t)(((cb + e) * TC_ROWS + lane_grp * 32 + lane) * 4)) : "memory");
#pragma unroll
          for (int e = 0; e < 8; e++) v[e] = __float_as_uint(__uint_as_float(v[e]) + t[e]);
        }
      }
      if (row < p.Mout) {
        const int col = n0 + cb;
        const float *rp = p.residual ? p.residual + (size_t)row * p.res_stride + p.res_off + col : nullptr;
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; e++) x[e] = __uint_as_float(v[e]);
        const bool full = col + 8 <= p.Cout;
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 8; e++)
            if (col + e < p.Cout) x[e] += __ldg(&p.bias[col + e]);
        }
        if (rp) {
          if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
            const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rp)), r1 = __ldg(reinterpret_cast<const float4 *>(rp) + 1);
            x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
            x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++)
              if (col + e < p.Cout) x[e] += __ldg(rp + e);
          }
        }
        if (p.out) {
          float *op = p.out + (size_t)row * p.out_stride + p.out_off + col;
          if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
            reinterpret_cast<float4 *>(op)[0] = make_float4(x[0], x[1], x[2], x[3]);
            reinterpret_cast<float4 *>(op)[1] = make_float4(x[4], x[5], x[6], x[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++)
              if (col + e < p.Cout) op[e] = x[e];
          }
        }
        if (p.pk) {
          // the consumer's BatchNorm(eval)+ReLU, then the fp16 hi/lo split: 8 channels -> 4 words hi + 4 words lo
          float y[8];
          bool big = false;
#pragma unroll
          for (int e = 0; e < 8; e++) {
            float t = x[e];
            if (col + e >= p.Cout) t = 0.f;  // padding channels stay exactly zero
            else if (p.pk_scale) t = fmaf(t, __ldg(&p.pk_scale[col + e]), __ldg(&p.pk_shift[col + e]));
            if (p.pk_relu) t = fmaxf(t, 0.f);
            big |= !(fabsf(t) <= 65504.f);
            y[e] = t;
          }
          if (big) *p.oflow = 1;
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const __half2 h = f2h2_sat(y[2 * q], y[2 * q + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = f2h2_sat((y[2 * q] - hf.x) * kLoScale, (y[2 * q + 1] - hf.y) * kLoScale);
            hw[q] = *reinterpret_cast<const uint32_t *>(&h);
            lw[q] = *reinterpret_cast<const uint32_t *>(&l);
          }
          const int ch = p.pk_coff + col;  // channel in the packed tensor (multiple of 8)
          uint32_t *dst = p.pk + (size_t)row * p.pk_stride + (ch >> 5) * 32 + ((ch & 31) >> 1);
          *reinterpret_cast<uint4 *>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          if (col + 8 == N && (N & 31) && p.pk_fill) {
            // N = Cout rounded to 16 ends in the middle of a 32-channel chunk: the consumer reads whole chunks, so the
            // upper half must be zero (not stale memory: 0 * NaN would poison the sums)
            const int ch2 = p.pk_coff + N;
            uint32_t *zp = p.pk + (size_t)row * p.pk_stride + (ch2 >> 5) * 32 + ((ch2 & 31) >> 1);
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            reinterpret_cast<uint4 *>(zp)[0] = zero; reinterpret_cast<uint4 *>(zp)[1] = zero;
            reinterpret_cast<uint4 *>(zp + 16)[0] = zero; reinterpret_cast<uint4 *>(zp + 16)[1] = zero;
          }
        }
      }
    }
  }
  if (split) {  // ranks > 0 keep their shared memory alive until rank 0 has read it
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// y = BatchNorm(eval)+ReLU(x) (or x when scale == nullptr) -> packed rows (see the header of this file). One thread per
// (row, word pair); channels past C are zero up to Cfill. Used where a tensor has no producing convolution to fuse into
// (network input, the concat half written by the encoder, gathered point rows).
__global__ void act_pack_kernel(const float *__restrict__ x, int x_stride, int x_off, const float *__restrict__ scale,
                                const float *__restrict__ shift, int relu, uint32_t *__restrict__ y, int y_stride, int y_coff,
                                int M, int C, int Cfill, int *__restrict__ oflow) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int wpr = Cfill >> 1;  // word pairs (2 channels) per row
  if (t >= (long long)M * wpr) return;
  const int row = (int)(t / wpr), pr = (int)(t % wpr);
  const int c = 2 * pr;
  float a = 0.f, b = 0.f;
  if (c < C) a = x[(size_t)row * x_stride + x_off + c];
  if (c + 1 < C) b = x[(size_t)row * x_stride + x_off + c + 1];
  if (scale) {
    if (c < C) a = fmaf(a, __ldg(&scale[c]), __ldg(&shift[c]));
    if (c + 1 < C) b = fmaf(b, __ldg(&scale[c + 1]), __ldg(&shift[c + 1]));
  }
  if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
  if (!(fabsf(a) <= 65504.f) || !(fabsf(b) <= 65504.f)) *oflow = 1;
  const __half2 h = f2h2_sat(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = f2h2_sat((a - hf.x) * kLoScale, (b - hf.y) * kLoScale);
  const int ch = y_coff + c;
  uint32_t *yr = y + (size_t)row * y_stride + (ch >> 5) * 32 + ((ch & 31) >> 1);
  yr[0] = *reinterpret_cast<const uint32_t *>(&h);
  yr[16] = *reinterpret_cast<const uint32_t *>(&l);
}

}  // namespace sgb

using namespace sgb;

namespace {

int *g_oflow[16] = {nullptr};  // per device: flag raised by the packing code when a value leaves the fp16 range

struct DevInfo { int sms; bool attr_set; bool ss_attr_set; };
DevInfo g_dev[16] = {};

int current_device(int *dev) {
  SGB_CUDA_CHECK(cudaGetDevice(dev));
  SGB_REQUIRE(*dev >= 0 && *dev < 16, SGB_ERR_RANGE, "device ordinal beyond 15");
  return SGB_OK;
}

int ensure_device(int dev) {
  if (!g_dev[dev].sms) {
    int sms = 0;
    SGB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev[dev].sms = sms;
  }
  if (!g_oflow[dev]) {
    SGB_CUDA_CHECK(cudaMalloc(&g_oflow[dev], 4));
    SGB_CUDA_CHECK(cudaMemset(g_oflow[dev], 0, 4));
  }
  return SGB_OK;
}

// Which kernel runs a problem (launch-by-launch A/B inside the real 150k-point step, profiles/r2_conv_launch_ab.txt): the
// persistent shared-memory-ring kernel (spconv_ss.cu) wins from 16 row tiles upwards when Cout >= 64 (levels 1-4 of the
// U-Net: -5 % ... -40 %) and on the big 32-channel levels; the register-gather kernel keeps the deep levels (a handful of
// row tiles: column split + split-K clusters) and the 32-channel tiny U-Net (a few hundred tiles, +10 % there).
bool use_ss_kernel(int K, int Mout, int Cin, int Cout, int sms) {
  int plan[5];
  if (!spconv_ss_plan(K, Mout, Cin, Cout, sms, plan)) return false;
  const int tiles = div_up(Mout, TC_ROWS);
  return tiles >= 16 && (Cout >= 64 || tiles >= 4 * sms);
}

}  // namespace

extern "C" {

// log2 of the factor the fp16 remainders (weights AND activations) are scaled by; the host weight packer uses the same.
int sgb_spconv_lo_shift(void) { return sgb::kLoShift; }

// Reads (and clears) the overflow flag raised by the packing code paths: blocking 4-byte read on `stream`.
int sgb_spconv_overflow(int *h_flag, void *stream) {
  SGB_REQUIRE(h_flag, SGB_ERR_ARG, "null flag");
  *h_flag = 0;
  int dev;
  int rc = current_device(&dev);
  if (rc) return rc;
  if (!g_oflow[dev]) return SGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SGB_CUDA_CHECK(cudaMemcpyAsync(h_flag, g_oflow[dev], 4, cudaMemcpyDeviceToHost, st));
  SGB_CUDA_CHECK(cudaStreamSynchronize(st));
  if (*h_flag) SGB_CUDA_CHECK(cudaMemsetAsync(g_oflow[dev], 0, 4, st));
  return SGB_OK;
}

int sgb_act_pack(const float *d_x, int x_stride, int x_off, const float *d_scale, const float *d_shift, int relu,
                 float *d_pk, int pk_stride, int pk_coff, int M, int C, int Cfill, void *stream) {
  if (M == 0 || C == 0) return SGB_OK;
  SGB_REQUIRE(d_x && d_pk && M > 0 && C > 0 && (d_scale == nullptr) == (d_shift == nullptr), SGB_ERR_ARG, "act_pack arguments");
  SGB_REQUIRE((pk_stride & 31) == 0 && (pk_coff & 1) == 0 && (Cfill & 1) == 0 && Cfill >= C && pk_stride >= pk_coff + Cfill,
              SGB_ERR_ARG, "act_pack: row stride multiple of 32 words, channel offset and fill width even, fill inside the row");
  int dev;
  int rc = current_device(&dev);
  if (rc) return rc;
  rc = ensure_device(dev);
  if (rc) return rc;
  long long tot = (long long)M * (Cfill / 2);
  act_pack_kernel<<<div_up(tot, 256), 256, 0, (cudaStream_t)stream>>>(d_x, x_stride, x_off, d_scale, d_shift, relu,
                                                                    (uint32_t *)d_pk, pk_stride, pk_coff, M, C, Cfill, g_oflow[dev]);
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}

// Packed weight size in floats for sgb_spconv_forward_tc: K * ceil(Cin/32) * 8 * N * 4 with N = Cout rounded to 16.
long long sgb_spconv_tc_packed_floats(int K, int Cin, int Cout) {
  int N = (Cout + 15) / 16 * 16;
  int nkc = (Cin + 31) / 32;
  return (long long)K * nkc * 4 * 2 * N * 4;  // fp16: 4 chunks of 8 halves, counted in 32-bit words
}

// Tile configuration chosen for a problem (pure function of the sizes; also what the launch uses): for tests and tools.
// out[0] = NT, [1] = column parts, [2] = split-K ranks, [3] = TMEM columns, [4] = pipeline stages, [5] = row tiles
int sgb_spconv_tc_plan(int K, int Mout, int Cin, int Cout, int has_map, int sms, int *out) {
  SGB_REQUIRE(out && K >= 1 && Mout >= 1 && Cin >= 1 && Cout >= 1, SGB_ERR_ARG, "spconv_tc_plan arguments");
  if (sms <= 0) sms = kNumSMs;
  const int N = (Cout + 15) / 16 * 16;
  // Column split: few row tiles (deep U-Net levels) would leave most SMs idle and make one CTA stream all the
  // weights, so N is cut into NT-column CTAs until the grid covers the machine (NT multiple of 16, >= 32).
  const int tiles = div_up(Mout, TC_ROWS);
  int NT = std::min(N, 128);  // [B_hi | B_lo] is one operand with 2*NT <= 256 columns
  if (N > 128) NT = (N / 2 + 15) / 16 * 16;
  while (NT > 32 && tiles * div_up(N, NT) < sms) {
    int nxt = (NT / 2 + 15) / 16 * 16;
    if (nxt >= NT) break;
    NT = nxt;
  }
  const size_t bstage = 2 * (size_t)NT * TC_KC * 2;  // weight ring stage (fp16 hi + lo) in shared memory
  // TMEM budget: accumulator [main | corrections] = 2*NT columns, then the A ring (32 columns per stage: hi + lo pairs;
  // an even number of stages because the pipeline moves in pairs). 256 columns (two CTAs per SM) when that leaves
  // >= 4 stages (two pairs, so gathers overlap the MMAs), else all 512. The weight ring has the same number of stages.
  const int acols0 = (2 * NT + 31) / 32 * 32;
  int tmem_cols = 256;
  int ns = ((256 - acols0) / 32) & ~1;
  if (ns < 4) { tmem_cols = 512; ns = ((512 - acols0) / 32) & ~1; }
  ns = std::min(ns, 6);
  const size_t ring_cap = (tmem_cols == 256) ? 64 * 1024 : 100 * 1024;  // one CTA per SM when it owns all of TMEM
  while (ns > 2 && (size_t)ns * bstage > ring_cap) ns -= 2;
  // split-K: when even the narrowest column split leaves most SMs idle, a cluster of S CTAs shares each tile's
  // iterations (S <= 8, the portable cluster size) and reduces through distributed shared memory.
  const int nkc = (Cin + TC_KC - 1) / TC_KC;
  const int ctas = tiles * div_up(N, NT);
  int S = 1;
  if (ctas * 2 <= sms && tmem_cols == 256 && (size_t)ns * bstage >= (size_t)NT * TC_ROWS * 4)
    S = std::max(1, std::min({8, 2 * sms / ctas, std::max(1, K * nkc / 4)}));
  out[0] = NT; out[1] = div_up(N, NT); out[2] = S; out[3] = tmem_cols; out[4] = ns; out[5] = tiles;
  (void)has_map;
  return SGB_OK;
}

// 1 = the persistent shared-memory-ring kernel would run this problem, 0 = the register-gather kernel (pure function).
int sgb_spconv_kernel_choice(int K, int Mout, int Cin, int Cout, int sms) {
  if (K < 1 || Mout < 1 || Cin < 1 || Cout < 1) return 0;
  return use_ss_kernel(K, Mout, Cin, Cout, sms > 0 ? sms : kNumSMs) ? 1 : 0;
}

int sgb_spconv_forward_tc(const float *d_in_pk, int in_stride, int Min, const int32_t *d_map, int K, int Mout,
                          const float *d_Wp, int Cin, int Cout, const float *d_residual, int res_stride, int res_off,
                          const float *d_bias, float *d_out, int out_stride, int out_off, float *d_pk_out, int pk_stride,
                          int pk_coff, const float *d_pk_scale, const float *d_pk_shift, int pk_relu, int pk_fill,
                          void *stream) {
  return sgb_spconv_forward_tc_ex(d_in_pk, in_stride, Min, d_map, K, Mout, d_Wp, Cin, Cout, d_residual, res_stride, res_off, d_bias,
                                  d_out, out_stride, out_off, d_pk_out, pk_stride, pk_coff, d_pk_scale, d_pk_shift, pk_relu, pk_fill,
                                  -1, stream);
}

int sgb_spconv_forward_tc_ex(const float *d_in_pk, int in_stride, int Min, const int32_t *d_map, int K, int Mout,
                             const float *d_Wp, int Cin, int Cout, const float *d_residual, int res_stride, int res_off,
                             const float *d_bias, float *d_out, int out_stride, int out_off, float *d_pk_out, int pk_stride,
                             int pk_coff, const float *d_pk_scale, const float *d_pk_shift, int pk_relu, int pk_fill, int kernel,
                             void *stream) {
  if (Mout == 0 || Cout == 0) return SGB_OK;
  SGB_REQUIRE(kernel >= -1 && kernel <= 1, SGB_ERR_ARG, "kernel: -1 (choose), 0 (register gather), 1 (shared-memory ring)");
  SGB_REQUIRE(d_in_pk && d_Wp && (d_out || d_pk_out) && K >= 1 && K <= 27 && Mout > 0 && Min > 0 && Cin > 0 && Cout > 0, SGB_ERR_ARG,
              "spconv_forward_tc arguments");
  SGB_REQUIRE(d_map || (K == 1 && Min >= Mout), SGB_ERR_ARG, "identity map requires K == 1");
  SGB_REQUIRE((in_stride & 31) == 0 && in_stride >= (Cin + 31) / 32 * 32, SGB_ERR_ARG, "packed input row stride (words, multiple of 32)");
  SGB_REQUIRE((((uintptr_t)d_in_pk) & 31) == 0, SGB_ERR_ARG, "packed input must be 32-byte aligned");
  SGB_REQUIRE(!d_out || out_stride >= out_off + Cout, SGB_ERR_ARG, "fp32 output row stride");
  SGB_REQUIRE(!d_pk_out || ((pk_stride & 31) == 0 && (pk_coff & 7) == 0 && pk_stride >= pk_coff + (Cout + 15) / 16 * 16), SGB_ERR_ARG,
              "packed output: row stride multiple of 32 words, channel offset multiple of 8");
  SGB_REQUIRE((d_pk_scale == nullptr) == (d_pk_shift == nullptr), SGB_ERR_ARG, "scale/shift must come together");
  const int N = (Cout + 15) / 16 * 16;
  SGB_REQUIRE(N <= 256 && Cin <= 512, SGB_ERR_RANGE, "spconv_forward_tc: Cout > 256 or Cin > 512 is not tiled");
  int dev;
  int rc = current_device(&dev);
  if (rc) return rc;
  rc = ensure_device(dev);
  if (rc) return rc;
  if (kernel == 1 || (kernel == -1 && use_ss_kernel(K, Mout, Cin, Cout, g_dev[dev].sms)))
    return spconv_ss_launch(d_in_pk, in_stride, d_map, K, Mout, d_Wp, Cin, Cout, d_residual, res_stride, res_off, d_bias, d_out,
                            out_stride, out_off, d_pk_out, pk_stride, pk_coff, d_pk_scale, d_pk_shift, pk_relu, pk_fill,
                            g_oflow[dev], g_dev[dev].sms, &g_dev[dev].ss_attr_set, (cudaStream_t)stream);
  int plan[6];
  rc = sgb_spconv_tc_plan(K, Mout, Cin, Cout, d_map != nullptr, g_dev[dev].sms, plan);
  if (rc) return rc;
  TcArgs p;
  p.in = d_in_pk; p.in_stride = in_stride; p.in_off = 0;
  p.map = d_map; p.K = K; p.Mout = Mout;
  p.Wp = d_Wp; p.Cin = Cin; p.N = N; p.Cout = Cout;
  p.residual = d_residual; p.res_stride = res_stride; p.res_off = res_off;
  p.bias = d_bias;
  p.out = d_out; p.out_stride = out_stride; p.out_off = out_off;
  p.pk = (uint32_t *)d_pk_out; p.pk_stride = pk_stride; p.pk_coff = pk_coff;
  p.pk_scale = d_pk_scale; p.pk_shift = d_pk_shift; p.pk_relu = pk_relu; p.pk_fill = pk_fill;
  p.oflow = g_oflow[dev];
  const int NT = plan[0], S = plan[2], tiles = plan[5];
  p.NT = NT;
  p.tmem_cols = plan[3];
  p.nstages = p.nbstages = plan[4];
  p.tmem_acols = (2 * NT + 31) / 32 * 32;
  p.ksplit = S;
  const size_t bstage = 2 * (size_t)NT * TC_KC * 2;
  const size_t map_bytes = d_map ? (size_t)K * TC_ROWS * 4 : 0;
  // at least 77 KB per CTA: never more than two CTAs per SM, so a third CTA cannot sit in tcgen05.alloc behind clusters
  // whose other ranks wait at barrier.cluster (each CTA holds >= 256 of the 512 TMEM columns until it exits)
  const size_t smem = std::max(bstage * p.nbstages + map_bytes + 1024, (size_t)77 * 1024);
  if (!g_dev[dev].attr_set) {
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(160 * 1024)));
    SGB_CUDA_CHECK(cudaFuncSetAttribute(spconv_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    g_dev[dev].attr_set = true;
  }
  dim3 grid(tiles, plan[1], S);
  if (S == 1) {
    spconv_tc_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = S;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SGB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, spconv_tc_kernel, p));
  }
  SGB_LAUNCH_CHECK();
  return SGB_OK;
}
}
