// Split modes (two 16-bit planes per operand), pass 1 as an N-HALF PIPELINE (included by dpn_tc.cu).
//
// pass1_ts_kernel runs a strict chain per tile: GEMM g (16 K-chunks x 3 MMAs) -> epilogue g (TMEM -> registers -> next A operand in
// TMEM) -> GEMM g+1 ...; the tensor pipe idles during every epilogue and the epilogue warps idle during every GEMM.  Here every GEMM
// is issued as two N = 128 halves with their own accumulators (ACC0 = TMEM columns [0,128), ACC1 = [128,256) - the same 256 columns,
// no extra tensor memory) and the K loop of a GEMM whose A operand lives in tensor memory is cut in two (kA = chunks 0..7 = features
// 0..127 of the previous layer, kB = chunks 8..15):
//
//     unit order of GEMM g :  (h0,kA)  (h1,kA)  (h0,kB) -> commit acc_ready[0]   (h1,kB) -> commit acc_ready[1]
//     epilogue g, half h   :  waits acc_ready[h], drains ACC_h, writes features [128h, 128h+128) of the next A operand, arrives epi_done[h]
//
//   * epilogue (g,h0) overwrites A[kA] while (h1,kB) still runs: legal, every reader of A[kA] - (h0,kA) and (h1,kA) - was issued
//     before the commit that released acc_ready[0];
//   * GEMM g+1 starts (h0,kA) as soon as epilogue (g,h0) is done (A[kA] written, ACC0 drained) - while epilogue (g,h1) is running;
//     (h1,kA), (h0,kB), (h1,kB) need epi_done[1].
// So the tensor pipe works during the h1 epilogue and the epilogue warps work during the (h1,kB) MMAs: per GEMM round
// e + 2q + max(q, e) instead of 4q + 2e (q = a quarter of the GEMM's MMA time, e = half an epilogue).
//
// GEMMs whose A operand comes from shared memory (G1: PE slices, G2b: PE6 slices) have no A hazard and run half by half
// (h0: all chunks, commit; h1: all chunks, commit), their A slices travel through the ring once per half.
//
// Ring: 27 slots of 8 KB.  A slot holds either one (K = 16 chunk, N half) piece of a weight image - images for this kernel are stored
// half-split, [chunk][half][plane][k-core][128 | 96 rows][16 B] - multicast to both CTAs of the cluster, or the [128 points x 16]
// two-plane slice of this tile's PE / PE6 tile.  208 KB of weights in flight against 144 KB in pass1_ts_kernel.
#pragma once

namespace np {
constexpr int NSLOT = 27;
constexpr int SLOT = 8192;
constexpr uint32_t COL_AH = 256, COL_AL = 384;
constexpr uint16_t MC_MASK = (uint16_t)((1u << CLUSTER) - 1);
constexpr int SMEM = NSLOT * SLOT + NVEC * H * 4 + TP * 4 * 4;
struct PipeNP {
  uint64_t full[NSLOT], empty[NSLOT], acc_ready[2], epi_done[2];
  uint32_t tmem_base;
};
static_assert(SMEM + (int)sizeof(PipeNP) + 1024 <= 227 * 1024, "pass1_np_kernel: ring + vectors + row sums + barriers must fit one SM's 227 KB");

struct NetImages {            // half-split images of one (sample, net) + this tile's feature tiles
  const uint8_t *W1, *W1T, *W2, *W2T, *Wd, *Wa, *WaT;
  const uint8_t *pe, *pe6;
};

// The static schedule of one net: both the producer and the MMA issuer walk it, so the ring order can never diverge.
//   v.unit(image, first chunk, chunks, rows per half, half, A tile in global memory (nullptr: A in tensor memory),
//          fresh accumulator, wait for epi_done[half] first, commit acc_ready[half] after)
template <class V>
__device__ __forceinline__ void ts_gemm(V& v, const uint8_t* img, const int Nh) {
  v.unit(img, 0, 8, Nh, 0, nullptr, true, true, false);
  v.unit(img, 0, 8, Nh, 1, nullptr, true, true, false);
  v.unit(img, 8, 8, Nh, 0, nullptr, false, false, true);
  v.unit(img, 8, 8, Nh, 1, nullptr, false, false, true);
}
template <class V>
__device__ __forceinline__ void walk_net(V& v, const NetImages& im, const int sweep, const bool first_net) {
  v.unit(im.W1, 0, 12, 128, 0, im.pe, true, !first_net, true);            // G1 = PE W1^T, half 0 / half 1
  v.unit(im.W1, 0, 12, 128, 1, im.pe, true, !first_net, true);
  v.unit(im.Wd, 0, 12, 128, 0, im.pe6, true, true, false);                // G2 = PE6 Wd^T (needs only the accumulator) + h1 W2^T
  v.unit(im.W2, 0, 8, 128, 0, nullptr, false, false, false);
  v.unit(im.Wd, 0, 12, 128, 1, im.pe6, true, true, false);
  v.unit(im.W2, 0, 8, 128, 1, nullptr, false, false, false);
  v.unit(im.W2, 8, 8, 128, 0, nullptr, false, false, true);
  v.unit(im.W2, 8, 8, 128, 1, nullptr, false, false, true);
  ts_gemm(v, im.Wa, 128);                                                 // G3 = c Wa^T
  if (sweep) {
    ts_gemm(v, im.WaT, 128);                                              // G4 = um Wa
    ts_gemm(v, im.W2T, 128);                                              // G5 = y W2
    if (sweep > 1) ts_gemm(v, im.W1T, 96);                                // G6 = qm W1 (N = 192: halves of 96)
  }
}
}  // namespace np

template <bool F16>
__global__ void __cluster_dims__(CLUSTER, 1, 1) __launch_bounds__(Geo<2>::THREADS, 1) pass1_np_kernel(const Work w, const int sweep) {
  constexpr int PL = 2;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ np::PipeNP pipe;
  uint8_t* ring = smem;
  float* svec = reinterpret_cast<float*>(smem + np::NSLOT * np::SLOT);
  float* rowsum = svec + NVEC * H;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int b = blockIdx.x / w.T, tl = blockIdx.x % w.T;
  const size_t g = blockIdx.x;
  if (tid == 0) {
    for (int s = 0; s < np::NSLOT; ++s) { mbar_init(&pipe.full[s], 1); mbar_init(&pipe.empty[s], CLUSTER); }
    for (int h = 0; h < 2; ++h) { mbar_init(&pipe.acc_ready[h], 1); mbar_init(&pipe.epi_done[h], Geo<PL>::ET); }
    fence_barrier_init();
  }
  if (warp == Geo<PL>::W_MMA) tmem_alloc(&pipe.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CLUSTER > 1) cluster_sync_all();
  const uint32_t tmem = pipe.tmem_base;

  auto images = [&](const int k) {
    const uint8_t* gen = w.img_gen_np + ((size_t)b * w.Kn + k) * (size_t)(PL * (2 * IMG_HC + 2 * IMG_HH));
    const uint8_t* sta = w.img_sta_np + (size_t)k * (size_t)(PL * (IMG_HC + 2 * IMG_HH));
    np::NetImages im;
    im.W1 = gen; im.W1T = gen + PL * IMG_HC; im.W2 = gen + PL * 2 * IMG_HC; im.W2T = gen + PL * (2 * IMG_HC + IMG_HH);
    im.Wd = sta; im.Wa = sta + PL * IMG_HC; im.WaT = sta + PL * (IMG_HC + IMG_HH);
    im.pe = w.pe_blob + g * Geo<PL>::BC; im.pe6 = w.pe6_blob + g * Geo<PL>::BC;
    return im;
  };

  if (warp == Geo<PL>::W_PROD) {
    // ---------------- producer: weight half-chunks (multicast slices) and this tile's PE / PE6 slices ----------------
    struct Prod {
      np::PipeNP* pp; uint8_t* ring; uint32_t rank; uint64_t pol; uint32_t s, ph; long long t_empty; bool timed;
      __device__ __forceinline__ void advance() { if (++s == np::NSLOT) { s = 0; ph ^= 1; } }
      __device__ __forceinline__ void unit(const uint8_t* img, const int c0, const int nch, const int Nh, const int half, const uint8_t* a_tile,
                                           bool, bool, bool) {
        const uint32_t hb = (uint32_t)Nh * 64u;                       // bytes of one (chunk, half): 2 planes x 2 k-cores x Nh rows x 16 B
        for (int c = c0; c < c0 + nch; ++c) {
          if (timed) mbar_wait_t(&pp->empty[s], ph ^ 1, t_empty); else mbar_wait(&pp->empty[s], ph ^ 1);
          if (elect_one()) {
            uint8_t* dst = ring + s * np::SLOT;
            const uint8_t* src = img + (size_t)(2 * c + half) * hb;
            mbar_arrive_expect_tx(&pp->full[s], hb);
            if (CLUSTER == 1) {
              bulk_g2s_hint(dst, src, hb, &pp->full[s], pol);
            } else {
              const uint32_t slice = hb / CLUSTER;
              bulk_g2s_mc_hint(dst + rank * slice, src + rank * slice, slice, &pp->full[s], np::MC_MASK, pol);
            }
          }
          advance();
          if (a_tile) {                                                // the K = 16 slice of the tile, both planes: 2 x 4 KB
            mbar_wait(&pp->empty[s], ph ^ 1);
            if (elect_one()) {
              uint8_t* dst = ring + s * np::SLOT;
              mbar_arrive_expect_tx(&pp->full[s], 2 * 4096);
#pragma unroll
              for (int p = 0; p < PL; ++p) bulk_g2s_hint(dst + p * 4096, a_tile + (size_t)p * BLOB_C + (size_t)c * 4096, 4096, &pp->full[s], pol);
            }
            advance();
          }
        }
      }
    } pr{&pipe, ring, cluster_ctarank(), l2_policy_evict_last(), 0u, 0u, 0ll, w.phase_dbg != nullptr};
    for (int k = 0; k < w.Kn; ++k) {
      const np::NetImages im = images(k);
      np::walk_net(pr, im, sweep, k == 0);
    }
    if (w.phase_dbg && lane == 0) atomicAdd((unsigned long long*)w.phase_dbg + 3, (unsigned long long)pr.t_empty);
  } else if (warp == Geo<PL>::W_MMA) {
    // ---------------- MMA issuer ----------------
    struct Iss {
      np::PipeNP* pp; uint32_t ring_addr, tmem; uint32_t s, ph; uint32_t we[2]; long long t_full, t_epi; bool timed, no_mma;
      __device__ __forceinline__ void advance() { if (++s == np::NSLOT) { s = 0; ph ^= 1; } }
      __device__ __forceinline__ void bwait(uint64_t* bar, uint32_t parity, long long& t) {
        if (timed) mbar_wait_t(bar, parity, t); else mbar_wait(bar, parity);
      }
      __device__ __forceinline__ void unit(const uint8_t*, const int c0, const int nch, const int Nh, const int half, const uint8_t* a_tile,
                                           const bool fresh, const bool wait, const bool commit) {
        if (wait) {                                                    // epilogue (g-1, half) has drained ACC_half and written its part of A
          bwait(&pp->epi_done[half], we[half] & 1, t_epi); ++we[half];
          tc_fence_after();
        }
        const uint32_t idesc = idesc_16(F16, Nh, 0, 0, 128);
        const uint32_t d = tmem + (uint32_t)half * 128u;
        const uint64_t b_base = smem_desc(ring_addr, (uint32_t)Nh * 16u, 128);
        const uint64_t a_base = smem_desc(ring_addr, CORE_STRIDE, 128);
        const uint32_t b_lo = ((uint32_t)Nh * 32u) >> 4;
        for (int c = c0; c < c0 + nch; ++c) {
          bwait(&pp->full[s], ph, t_full);
          const uint32_t sW = s;
          advance();
          uint32_t sA = 0;
          if (a_tile) { bwait(&pp->full[s], ph, t_full); sA = s; advance(); }
          tc_fence_after();
          const uint64_t bd = b_base + sW * (uint32_t)(np::SLOT >> 4), bl = bd + b_lo;
          const uint32_t first = (!fresh || c > c0) ? 1u : 0u;
          if (elect_one()) {
            if (!a_tile) {                                             // lo*hi + hi*lo + hi*hi, A planes from tensor memory
              if (!no_mma) {
                mma_ts(d, tmem + np::COL_AL + 8 * c, bd, idesc, first);
                mma_ts(d, tmem + np::COL_AH + 8 * c, bl, idesc, 1u);
                mma_ts(d, tmem + np::COL_AH + 8 * c, bd, idesc, 1u);
              }
              if (CLUSTER == 1) mma_commit(&pp->empty[sW]); else mma_commit_mc(&pp->empty[sW], np::MC_MASK);
            } else {
              const uint64_t ad = a_base + sA * (uint32_t)(np::SLOT >> 4), al = ad + (4096 >> 4);
              if (!no_mma) {
                mma_bf16(d, al, bd, idesc, first);
                mma_f16_c<REUSE_A ? A_FILL : A_DISCARD>(d, ad, bl, idesc, 1u);
                mma_f16_c<REUSE_A ? A_LAST : A_DISCARD>(d, ad, bd, idesc, 1u);
              }
              if (CLUSTER == 1) { mma_commit(&pp->empty[sW]); mma_commit(&pp->empty[sA]); }
              else { mma_commit_mc(&pp->empty[sW], np::MC_MASK); mma_commit_mc(&pp->empty[sA], np::MC_MASK); }
            }
          }
        }
        if (commit && elect_one()) mma_commit(&pp->acc_ready[half]);
      }
    } is{&pipe, smem_u32(ring), tmem, 0u, 0u, {0u, 0u}, 0ll, 0ll, w.phase_dbg != nullptr, DPN_DBG(w, 1)};
    const long long t_begin = clock64();
    for (int k = 0; k < w.Kn; ++k) {
      const np::NetImages im = images(k);
      np::walk_net(is, im, sweep, k == 0);
    }
    if (w.phase_dbg && lane == 0) {
      atomicAdd((unsigned long long*)w.phase_dbg + 0, (unsigned long long)(clock64() - t_begin));
      atomicAdd((unsigned long long*)w.phase_dbg + 1, (unsigned long long)is.t_full);
      atomicAdd((unsigned long long*)w.phase_dbg + 2, (unsigned long long)is.t_epi);
    }
  } else if (warp < Geo<PL>::EW) {
    // ---------------- epilogue: thread = (point r, 64-column group `sub` of the current N half) ----------------
    const int sub = warp >> 2, r = (warp & 3) * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int p_local = tl * TP + r;
    const bool valid = p_local < w.P;
    const size_t q = (size_t)b * w.N + w.p0 + p_local;
    const size_t row = g * TP + r;
    const float* pet = w.pet + g * (size_t)(C * TP) + r;
    const uint64_t pol_keep = l2_policy_evict_last();
    uint32_t ar[2] = {0u, 0u};
    long long t_acc = 0;
    const long long t_begin = clock64();
    const bool timed = w.phase_dbg != nullptr, skip_epi = DPN_DBG(w, 2), skip_st = DPN_DBG(w, 4);
    // 32 columns (block cg of the 256) of this row: split into the two planes once, then -> workspace tile (if any) and / or the next A operand
    auto emit = [&](const int cg, const float (&v)[32], uint8_t* blob, const bool to_a) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        uint4 pq[PL];
        split8<PL, F16>(v + qd * 8, pq);
        if (blob && !skip_st) {
          const uint32_t off = gp_off(32, r, cg * 4 + qd);
          __stcs(reinterpret_cast<uint4*>(blob + off), pq[0]);
          __stcs(reinterpret_cast<uint4*>(blob + BLOB_H + off), pq[1]);
        }
        hi[qd * 4 + 0] = pq[0].x; hi[qd * 4 + 1] = pq[0].y; hi[qd * 4 + 2] = pq[0].z; hi[qd * 4 + 3] = pq[0].w;
        lo[qd * 4 + 0] = pq[1].x; lo[qd * 4 + 1] = pq[1].y; lo[qd * 4 + 2] = pq[1].z; lo[qd * 4 + 3] = pq[1].w;
      }
      if (to_a) {
        tmem_st16(lane_base + np::COL_AH + cg * 16, hi);
        tmem_st16(lane_base + np::COL_AL + cg * 16, lo);
      }
    };
    auto acc_wait = [&](const int h) {
      if (timed) mbar_wait_t(&pipe.acc_ready[h], ar[h] & 1, t_acc); else mbar_wait(&pipe.acc_ready[h], ar[h] & 1);
      ++ar[h]; tc_fence_after();
    };
    auto done = [&](const int h) { tmem_st_wait(); tc_fence_before(); mbar_arrive(&pipe.epi_done[h]); };
    if (sub == 0) { rowsum[r * 4 + 0] = 0.f; rowsum[r * 4 + 1] = 0.f; rowsum[r * 4 + 2] = 0.f; rowsum[r * 4 + 3] = 0.f; }
    for (int k = 0; k < w.Kn; ++k) {
      uint8_t* nt = net_tile<PL>(w, b, k, tl);
      epi_bar<PL>();
      if (tid < H) load_vectors(svec, w, b, k, tid);
      epi_bar<PL>();
      float i1 = 1.f, i2 = 1.f, i3 = 1.f, i4 = 1.f, i5 = 1.f, i6 = 1.f, sH1 = 1.f, sC = 1.f, sG = 1.f, sUM = 1.f, sY = 1.f;
      if (F16) {
        const NetScales t = w.sc[b * w.Kn + k];
        i1 = 1.f / (S_PE * t.sW1); i2 = 1.f / (t.sH1 * t.sW2); i3 = 1.f / (t.sC * t.sWa); i4 = 1.f / (t.sUM * t.sWa);
        i5 = t.sQ / (t.sY * t.sW2); i6 = 1.f / (t.sQ * t.sW1);
        sH1 = t.sH1; sC = t.sC; sG = t.sG; sUM = t.sUM; sY = t.sY;
      }
      uint32_t m1w[4] = {0u, 0u, 0u, 0u};                            // ReLU mask of a1 for this thread's four 32-column blocks (half, block)
      // ---- epilogue 1: h1 = relu(a1 + b1) ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
#pragma unroll 1
        for (int cb = 0; cb < (skip_epi ? 0 : 2); ++cb) {
          const int cg = 4 * h + 2 * sub + cb;
          float v[32];
          tmem_ld32(lane_base + cg * 32, v);
          uint32_t bits = 0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bv = *reinterpret_cast<const float4*>(svec + V_B1 * H + cg * 32 + j4 * 4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = F16 ? fmaf(v[j4 * 4 + e], i1, bb[e]) : v[j4 * 4 + e] + bb[e];
              bits |= (a > 0.f ? 1u : 0u) << (j4 * 4 + e);
              v[j4 * 4 + e] = F16 ? fmaxf(a, 0.f) * sH1 : fmaxf(a, 0.f);
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) m1w[i] = (h * 2 + cb == i) ? bits : m1w[i];
          emit(cg, v, sweep ? blob_h<PL>(nt, B_H1) : nullptr, true);
        }
        done(h);
      }
      // ---- epilogue 2: c = acc + (b2 + bd + e);  oc = 2wo.c ----
      float os0 = 0.f, os1 = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
#pragma unroll 1
        for (int cb = 0; cb < (skip_epi ? 0 : 2); ++cb) {
          const int cg = 4 * h + 2 * sub + cb;
          float v[32];
          tmem_ld32(lane_base + cg * 32, v);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 bv = *reinterpret_cast<const float4*>(svec + V_BSUM * H + cg * 32 + j4 * 4);
            const float4 wv = *reinterpret_cast<const float4*>(svec + V_WO2 * H + cg * 32 + j4 * 4);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float cc = F16 ? fmaf(v[j4 * 4 + e], i2, bb[e]) : v[j4 * 4 + e] + bb[e];
              if (e & 1) os1 = fmaf(ww[e], cc, os1); else os0 = fmaf(ww[e], cc, os0);
              v[j4 * 4 + e] = F16 ? cc * sC : cc;
            }
          }
          emit(cg, v, sweep ? blob_h<PL>(nt, B_CC) : nullptr, true);
        }
        done(h);
      }
      // ---- epilogue 3: g = relu(a3 + ba);  o = oc + u.g + cst + ref;  um = u*[a3>0] ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
#pragma unroll 1
        for (int cb = 0; cb < (skip_epi ? 0 : 2); ++cb) {
          const int cg = 4 * h + 2 * sub + cb;
          float v[32];
          tmem_ld32(lane_base + cg * 32, v);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            float g8[8];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const float4 bv = *reinterpret_cast<const float4*>(svec + V_BA * H + cg * 32 + qd * 8 + h2 * 4);
              const float4 uv = *reinterpret_cast<const float4*>(svec + V_U * H + cg * 32 + qd * 8 + h2 * 4);
              const float bb[4] = {bv.x, bv.y, bv.z, bv.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = qd * 8 + h2 * 4 + e;
                const float a = F16 ? fmaf(v[j], i3, bb[e]) : v[j] + bb[e];
                const float gg = fmaxf(a, 0.f);
                if (e & 1) os1 = fmaf(uu[e], gg, os1); else os0 = fmaf(uu[e], gg, os0);
                g8[h2 * 4 + e] = F16 ? gg * sG : gg;
                v[j] = a > 0.f ? (F16 ? uu[e] * sUM : uu[e]) : 0.f;          // um replaces the accumulator value in place
              }
            }
            if (sweep && !skip_st) {
              uint4 pq[PL];
              split8<PL, F16>(g8, pq);
              const uint32_t off = gp_off(32, r, cg * 4 + qd);
              __stcs(reinterpret_cast<uint4*>(blob_h<PL>(nt, B_GG) + off), pq[0]);
              __stcs(reinterpret_cast<uint4*>(blob_h<PL>(nt, B_GG) + BLOB_H + off), pq[1]);
            }
          }
          if (sweep) emit(cg, v, blob_h<PL>(nt, B_UM), true);
        }
        done(h);
      }
      atomicAdd(rowsum + r * 4, os0 + os1);                          // the two column groups of a row meet in shared memory
      epi_bar<PL>();
      if (sub == 0) {
        const float ref = w.ref ? __ldg(w.ref + q * w.Kn + k) : __ldg(w.coord_data + q * 6 + k);
        if (valid) w.o[row * w.Kn + k] = rowsum[r * 4] + __ldg(w.cst + k) + ref;
        rowsum[r * 4] = 0.f;
      }
      if (!sweep) continue;
      // ---- epilogue 4: y = acc + 2wo ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
#pragma unroll 1
        for (int cb = 0; cb < (skip_epi ? 0 : 2); ++cb) {
          const int cg = 4 * h + 2 * sub + cb;
          float v[32];
          tmem_ld32(lane_base + cg * 32, v);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 wv = *reinterpret_cast<const float4*>(svec + V_WO2 * H + cg * 32 + j4 * 4);
            const float ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j4 * 4 + e] = F16 ? fmaf(v[j4 * 4 + e], i4, ww[e]) * sY : v[j4 * 4 + e] + ww[e];
          }
          emit(cg, v, blob_h<PL>(nt, B_YT), true);
        }
        done(h);
      }
      // ---- epilogue 5: qm = acc * m1 ----
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
#pragma unroll 1
        for (int cb = 0; cb < (skip_epi ? 0 : 2); ++cb) {
          const int cg = 4 * h + 2 * sub + cb;
          float v[32];
          tmem_ld32(lane_base + cg * 32, v);
          uint32_t bits = 0u;
#pragma unroll
          for (int i = 0; i < 4; ++i) bits = (h * 2 + cb == i) ? m1w[i] : bits;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((bits >> j) & 1u) ? (F16 ? v[j] * i5 : v[j]) : 0.f;
          emit(cg, v, blob_h<PL>(nt, B_QM), sweep > 1);
        }
        done(h);
      }
      if (sweep < 2) continue;
      // ---- epilogue 6: do/dz_c = sum_j jin_j dPE_j  (j % 3 == c).  N = 192 as two halves of 96 columns (ACC0 / ACC1); this thread
      //      takes the 48-column group `sub` of each: features 96 h + 48 sub .. + 47 (a multiple of 6: the sin / cos pattern is static)
      float dz[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        acc_wait(h);
        if (skip_epi) { done(h); continue; }
        const int fb = 96 * h + 48 * sub;
        const uint32_t a6 = lane_base + (uint32_t)(128 * h + 48 * sub);
        {
          float pp[32], v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) pp[j] = ldg_f32_hint(pet + (size_t)(fb + DPE_PARTNER(j)) * TP, pol_keep);
          tmem_ld32(a6, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) dz[j % 3] = fmaf(DPE_SIGN(j) * w.band[16 * h + 8 * sub + j / 6] * v[j], pp[j], dz[j % 3]);
        }
        {
          float pp[16], v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pp[j] = ldg_f32_hint(pet + (size_t)(fb + DPE_PARTNER(32 + j)) * TP, pol_keep);
          tmem_ld16(a6 + 32, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) dz[(32 + j) % 3] = fmaf(DPE_SIGN(32 + j) * w.band[16 * h + 8 * sub + (32 + j) / 6] * v[j], pp[j], dz[(32 + j) % 3]);
        }
        done(h);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) atomicAdd(rowsum + r * 4 + 1 + c, F16 ? dz[c] * i6 : dz[c]);
      epi_bar<PL>();
      if (sub == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          if (valid) w.od[(row * w.Kn + k) * 3 + c] = rowsum[r * 4 + 1 + c];
          rowsum[r * 4 + 1 + c] = 0.f;
        }
      }
    }
    if (w.phase_dbg && tid == 0) {
      const long long tot = clock64() - t_begin;
      atomicAdd((unsigned long long*)w.phase_dbg + 4, (unsigned long long)tot);
      atomicAdd((unsigned long long*)w.phase_dbg + 5, (unsigned long long)t_acc);
      atomicAdd((unsigned long long*)w.phase_dbg + 7, (unsigned long long)(tot - t_acc));
      atomicAdd((unsigned long long*)w.phase_dbg + 6, 1ull);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync_all();
  if (warp == Geo<PL>::W_MMA) tmem_dealloc(tmem, 512);
}
