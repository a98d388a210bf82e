// MappedGrid 3-D FULL profile as FOUR light kernels (round 2): conserved metrics, xi faces, eta faces, zeta faces
// (+ the cell arrays).  Replaces the launch of /root/reference/src/grids/unified_grids/metrics.jl:2079-2334 (cell fill
// + three face fills) for separable term-list mappings; arithmetic as in cg_mapped.cuh.
//
// Why four kernels.  With the warp-uniform eta / zeta products factored out (k_mapped_metrics3d_uni, cg_mapped.cuh)
// the fused kernel does ~1400 FP64 operations per cell but still needs 16.9 k cycles per warp step at 8 warps/SM:
// every short loop (two terms per coordinate, a table row per pair) exposes its L1 / shared-memory round trip, and
// 223-243 registers leave two warps per scheduler to hide them (ncu: FP64 pipe 34 %, 1.4 long-scoreboard + 1.5 wait
// stalls per issue; software pipelining of the loads and 12 warps with spills both measured slower).  The work of a
// step splits cleanly, though:
//   k_mapped_cons_uni      27 conserved entries: 15 multiply-adds per pair, no Jacobian        216 B/cell written
//   k_mapped_face_uni<0>   xi faces: F, dF/dxi, d2F/dxi2 -> L / R states, +xi neighbour by shuffle  160 B/cell
//   k_mapped_face_uni<1>   eta faces, marching along j: the +eta neighbour is the next step     160 B/cell
//   k_mapped_face_uni<2>   zeta faces, marching along k, + the cell arrays (F, G, J are there)  320 B/cell
// Each holds one 3x3 accumulation set and one L / R evaluation at a time (no carried Jacobian derivatives, no
// recomputed +eta neighbour: 1140 FP64 operations per cell against 1415), fits 96-128 registers, and runs 16+ one-warp
// blocks per SM, which is what hides the latencies.  The xi rows a lane reads are the same for every step and for every
// row, so blocks are ordered row-fastest and share them in L1.
#pragma once
#include <type_traits>

#include "cg_mapped.cuh"

namespace cg {

#ifndef CG_SPLIT_CONS_MINB
#define CG_SPLIT_CONS_MINB 16
#endif
#ifndef CG_SPLIT_FACE_MINB
#define CG_SPLIT_FACE_MINB 12
#endif

// ---------------------------------------------------------------------------
// conserved metrics of all three face axes (lanes along xi, rows j, marching along k, TWO planes per pass)
//
// The kernel is bound by the load / store unit, not by arithmetic (180 multiply-adds per cell): per pair a lane needs
// its six xi values and the 15 warp-uniform coefficients (8 broadcast LDS.128 = 16 wavefronts).  Measured with one
// plane per pass and AoS xi rows: 10.6 ms at 512^3, LSU wavefronts 90 % of peak, DRAM 36 %.  Hence
//   * the xi rows come from the transposed, 128-byte aligned copy `tabx` (one coalesced 256-byte segment per value),
//   * a pass over the pairs serves the planes m and m + 1 (the xi values are the same), and
//   * the sums run column by column (the pairs of a column only touch the column's 9 entries per plane), so 18
//     accumulators are live instead of 54 and the kernel keeps 12+ one-warp blocks per SM.
// ---------------------------------------------------------------------------
constexpr int CONS_KS = 2;  // planes per pass
template <class T>
__global__ void __launch_bounds__(32, CG_SPLIT_CONS_MINB) k_mapped_cons_uni(Dims D, MPairs PR, const T* __restrict__ tab,
                                                                             const T* __restrict__ tabx, int64_t L, int64_t Lx,
                                                                             MetricOut<T> O, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const int64_t j = (int64_t)blockIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * 32;
  const int nvalid = (int)(nw0 - i0 < 32 ? nw0 - i0 : 32);
  const int64_t ic = i0 + lane < nw0 ? i0 + lane : nw0 - 1;  // lanes past the row end repeat the last cell (never sent)
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  const int np = PR.n;
  T* const b9 = reinterpret_cast<T*>(cg_dyn_smem);           // [plane][face] ConservedMetric staging buffers
  T* const pairU = b9 + CONS_KS * 3 * stage_len<T, 9>;       // [plane][np][UNI_PAIR]
  const T* const xcol = tabx + ic;  // + (p * OP_COUNT + op) * Lx : operator value `op` of pair p at this lane's cell
  const T* const yrow = tab + (L + j) * OP_COUNT;
  const T* const zrow = tab + 2 * L * OP_COUNT;
  const int64_t pstride = 3 * L * OP_COUNT;
  // coefficient tables of the planes mz, mz + 1: work item w = plane * np + pair (see k_mapped_metrics3d_uni for the
  // 15 coefficients); zpre = the zeta row of the work item `lane`, loaded earlier
  auto pair_uniform = [&](int64_t mz, const T (&zpre)[OP_COUNT]) {
    for (int w = lane; w < CONS_KS * np; w += 32) {
      const int pl = w >= np ? 1 : 0, p = w - pl * np;
      T y[OP_COUNT], z[OP_COUNT];
      ldg_vec<T, OP_COUNT>(yrow + p * pstride, y);
      if (w == lane) {
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) z[o] = zpre[o];
      } else {
        const int64_t mm = mz + pl < nw2 ? mz + pl : nw2 - 1;
        ldg_vec<T, OP_COUNT>(zrow + p * pstride + mm * OP_COUNT, z);
      }
      const T C = T(PR.coef[p]);
#pragma unroll
      for (int o = 0; o < OP_COUNT; ++o) z[o] *= C;
      T u[UNI_PAIR];
      u[0] = y[OP_V1] * z[OP_D0] - y[OP_D0] * z[OP_V1];
      u[1] = y[OP_V0] * z[OP_V1];
      u[2] = -(z[OP_V0] * y[OP_V1]);
      u[3] = -(y[OP_V0] * z[OP_D0]);
      u[4] = z[OP_V0] * y[OP_D0];
      u[5] = y[OP_E0] * z[OP_V1];
      u[6] = -(z[OP_V0] * y[OP_E1]);
      u[7] = -(z[OP_E0] * y[OP_V1]);
      u[8] = y[OP_V0] * z[OP_E1];
      u[9] = -(y[OP_E0] * z[OP_D0]);
      u[10] = z[OP_V0] * y[OP_FD0];
      u[11] = z[OP_E0] * y[OP_D0];
      u[12] = -(y[OP_V0] * z[OP_FD0]);
      u[13] = y[OP_E1] * z[OP_D0] - y[OP_FD0] * z[OP_V1];
      u[14] = z[OP_FD0] * y[OP_V1] - z[OP_E1] * y[OP_D0];
      u[15] = T(0);
      sts_vec<T, UNI_PAIR>(pairU + w * UNI_PAIR, u);
    }
  };
  auto load_zpair = [&](int64_t mz, T (&z)[OP_COUNT]) {
    const int w = lane < CONS_KS * np ? lane : CONS_KS * np - 1;
    const int pl = w >= np ? 1 : 0, p = w - pl * np;
    const int64_t mm = mz + pl < nw2 ? mz + pl : nw2 - 1;
    ldg_vec<T, OP_COUNT>(zrow + p * pstride + mm * OP_COUNT, z);
  };
  {
    T zp[OP_COUNT];
    load_zpair(k0, zp);
    pair_uniform(k0, zp);
    __syncwarp();
  }
  for (int64_t m = k0; m < k1; m += CONS_KS) {
    const int64_t cl0 = i0 + nw0 * (j + nw1 * m);
    const bool two = m + 1 < k1;              // the second plane of the pass exists
    const bool more = m + CONS_KS < k1;
    T zp[OP_COUNT];  // the zeta rows of the next pass: requested now, multiplied after the sums
    load_zpair(more ? m + CONS_KS : m, zp);
    bulk_wait_read();  // the TMA unit has finished reading last pass's staging buffers (lanes > 0 own no groups)
    __syncwarp();
    // staging slots of this lane: [plane][face], element = 3x3 column-major (row + 3 * col)
    Stage<T, 9> S00(O.face_cons[0], cl0, b9), S01(O.face_cons[1], cl0, b9 + stage_len<T, 9>),
        S02(O.face_cons[2], cl0, b9 + 2 * stage_len<T, 9>);
    const int64_t cl1 = cl0 + nw0 * nw1;
    Stage<T, 9> S10(O.face_cons[0], cl1, b9 + 3 * stage_len<T, 9>), S11(O.face_cons[1], cl1, b9 + 4 * stage_len<T, 9>),
        S12(O.face_cons[2], cl1, b9 + 5 * stage_len<T, 9>);
    T* const o00 = S00.slot(lane), * const o01 = S01.slot(lane), * const o02 = S02.slot(lane);
    T* const o10 = S10.slot(lane), * const o11 = S11.slot(lane), * const o12 = S12.slot(lane);
#pragma unroll
    for (int cc = 0; cc < 3; ++cc) {
      T A[CONS_KS][3][3];  // [plane][face][row] of column cc
#pragma unroll
      for (int pl = 0; pl < CONS_KS; ++pl)
#pragma unroll
        for (int f = 0; f < 3; ++f)
#pragma unroll
          for (int rr = 0; rr < 3; ++rr) A[pl][f][rr] = T(0);
#pragma unroll(2)
      for (int p = PR.cstart[cc]; p < PR.cstart[cc + 1]; ++p) {
        T x[OP_COUNT];
        const T* xp = xcol + (int64_t)p * OP_COUNT * Lx;
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) x[o] = __ldg(xp + o * Lx);  // coalesced: transposed xi rows
#pragma unroll
        for (int pl = 0; pl < CONS_KS; ++pl) {
          T u[UNI_PAIR];
          lds_vec<T, UNI_PAIR>(pairU + (pl * np + p) * UNI_PAIR, u);
          A[pl][0][0] += x[OP_E0] * u[0];
          A[pl][0][1] += x[OP_FD0] * u[1] + x[OP_E1] * u[3];
          A[pl][0][2] += x[OP_FD0] * u[2] + x[OP_E1] * u[4];
          A[pl][1][1] += x[OP_D0] * u[5] + x[OP_V1] * u[9];
          A[pl][1][2] += x[OP_D0] * u[6] + x[OP_V1] * u[10];
          A[pl][1][0] += x[OP_V0] * u[13];
          A[pl][2][2] += x[OP_D0] * u[7] + x[OP_V1] * u[11];
          A[pl][2][1] += x[OP_D0] * u[8] + x[OP_V1] * u[12];
          A[pl][2][0] += x[OP_V0] * u[14];
        }
      }
      // every lane stages (no branch); slots >= nvalid and the second plane of an odd tail are never sent
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        o00[rr + 3 * cc] = A[0][0][rr]; o01[rr + 3 * cc] = A[0][1][rr]; o02[rr + 3 * cc] = A[0][2][rr];
        o10[rr + 3 * cc] = A[1][0][rr]; o11[rr + 3 * cc] = A[1][1][rr]; o12[rr + 3 * cc] = A[1][2][rr];
      }
    }
    __syncwarp();  // every lane has read pairU
    if (more) pair_uniform(m + CONS_KS, zp);
    fence_async_smem();
    __syncwarp();
    S00.issue(nvalid, lane);
    S01.issue(nvalid, lane);
    S02.issue(nvalid, lane);
    if (two) {
      S10.issue(nvalid, lane);
      S11.issue(nvalid, lane);
      S12.issue(nvalid, lane);
    }
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// ---------------------------------------------------------------------------
// face forward / inverse metrics of ONE face axis AX (+ the cell arrays when AX == 2).
//   AX = 0 : rows j, marching along k; the +xi neighbour's R state comes by warp shuffle (lane 31 = cell i0 + 31)
//   AX = 1 : rows k, marching along j; AX = 2 : rows j, marching along k; the +AX neighbour is the cell of the NEXT
//            step: each step evaluates the cell m + 1 (F, G, J, L, R), averages its R state with the carried L state
//            of the cell m and stores the face m + 1/2
// Uniform products of a step, one row of UNI_TERM values per term (flat, coordinate-ordered index s), with
// P(a, b) = coef f_R^(a)(row) f_M^(b)(march position), R / M = the row / marching axis:
//   AX = 0 : [P00 P10 P01 -]                      F, dF/dxi, d2F/dxi2 = xi-factor orders (1,0,0 | 2,1,1 | 3,2,2) x these
//   AX = M : [P00 P10 P01 P11 P02 P12 P03 -]      F | dF/dxi_M | d2F/dxi_M^2
// ---------------------------------------------------------------------------
// RX = true (mappings with at most FACE_NT terms): the lane's xi factor rows — four derivative orders per term, the same
// for every step of the march — live in registers instead of being re-read from L1 in every step.
constexpr int FACE_NT = 8;
template <class T, int SCH, int AX, bool RX>
__global__ void __launch_bounds__(32, CG_SPLIT_FACE_MINB) k_mapped_face_uni(Dims D, MCoordTerms CT, const T* __restrict__ ftab,
                                                                             int64_t L, MetricOut<T> O, int nt, int chunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  constexpr int MA = AX == 1 ? 1 : 2;  // marching axis
  constexpr int RA = AX == 1 ? 2 : 1;  // row axis
  constexpr int OUT = AX == 0 ? 31 : 32;
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  const int64_t L1 = L + 1;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1];
  const int64_t row = (int64_t)blockIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * OUT;
  const int nvalid = (int)(nw0 - i0 < OUT ? nw0 - i0 : OUT);
  // lanes past the row end (and lane 31 of the xi kernel) evaluate up to the cell nw0: the factor tables hold nw + 1
  const int64_t i = i0 + lane < nw0 ? i0 + lane : nw0;
  const int64_t m0 = (int64_t)blockIdx.z * chunk;
  const int64_t m1 = m0 + chunk < D.nw[MA] ? m0 + chunk : D.nw[MA];
  T* const b10 = reinterpret_cast<T*>(cg_dyn_smem);  // two Metric staging buffers
  T* const termU = b10 + 2 * stage_len<T, 10>;       // [nt][UNI_TERM]
  const int64_t tstride = 3 * L1 * 4;
  const T* const fx = ftab + i * 4;  // + m * tstride : xi factor row (four derivative orders) of term m
  // the term this lane multiplies in the uniform phase: its row-axis factors stay in registers
  const int su = lane < nt ? lane : nt - 1;
  const T* const ubase = ftab + (int64_t)CT.flat[su] * tstride;
  T yr[4];
  ldg_vec<T, 4>(ubase + (RA * L1 + row) * 4, yr);
  {
    const T c = T(CT.coef[su]);
#pragma unroll
    for (int o = 0; o < 4; ++o) yr[o] *= c;
  }
  auto load_z = [&](int64_t mm, T (&z)[4]) { ldg_vec<T, 4>(ubase + (MA * L1 + mm) * 4, z); };
  auto term_uniform = [&](const T (&z)[4]) {
    if (lane < nt) {
      T u[UNI_TERM];
      u[0] = yr[0] * z[0]; u[1] = yr[1] * z[0]; u[2] = yr[0] * z[1];
      if (AX == 0) {
        u[3] = T(0);
        T v[4] = {u[0], u[1], u[2], u[3]};
        sts_vec<T, 4>(termU + lane * UNI_TERM, v);
      } else {
        u[3] = yr[1] * z[1]; u[4] = yr[0] * z[2]; u[5] = yr[1] * z[2]; u[6] = yr[0] * z[3]; u[7] = T(0);
        sts_vec<T, UNI_TERM>(termU + lane * UNI_TERM, u);
      }
    }
  };
  T ar[RX ? FACE_NT : 1][4];
  const int cn0 = CT.cnt[0], cn1 = cn0 + CT.cnt[1];
  if (RX) {
#pragma unroll
    for (int sx = 0; sx < FACE_NT; ++sx) ldg_vec<T, 4>(fx + (int64_t)CT.flat[sx < nt ? sx : nt - 1] * tstride, ar[RX ? sx : 0]);
  }
  // one term's contribution to the row q of F, dF/dxi_AX, d2F/dxi_AX^2 (columns 0 / RA / MA)
  auto term_acc = [&](const T (&a)[4], int sx, T (&F)[3], T (&Fp)[3], T (&Fpp)[3]) {
    if (AX == 0) {
      T u[4];
      lds_vec<T, 4>(termU + sx * UNI_TERM, u);
      F[0] += a[1] * u[0]; F[RA] += a[0] * u[1]; F[MA] += a[0] * u[2];
      Fp[0] += a[2] * u[0]; Fp[RA] += a[1] * u[1]; Fp[MA] += a[1] * u[2];
      Fpp[0] += a[3] * u[0]; Fpp[RA] += a[2] * u[1]; Fpp[MA] += a[2] * u[2];
    } else {
      T u[UNI_TERM];
      lds_vec<T, UNI_TERM>(termU + sx * UNI_TERM, u);
      F[0] += a[1] * u[0]; F[RA] += a[0] * u[1]; F[MA] += a[0] * u[2];
      Fp[0] += a[1] * u[2]; Fp[RA] += a[0] * u[3]; Fp[MA] += a[0] * u[4];
      Fpp[0] += a[1] * u[4]; Fpp[RA] += a[0] * u[5]; Fpp[MA] += a[0] * u[6];
    }
  };
  // F, dF/dxi_AX, d2F/dxi_AX^2 at the position the uniform table was built for
  auto jacobian = [&](T (&F)[3][3], T (&Fp)[3][3], T (&Fpp)[3][3]) {
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = Fp[q][d] = Fpp[q][d] = T(0);
    if (RX) {
#pragma unroll
      for (int sx = 0; sx < FACE_NT; ++sx) {
        if (sx < nt) {  // the coordinate a term belongs to is warp-uniform
          if (sx < cn0) term_acc(ar[RX ? sx : 0], sx, F[0], Fp[0], Fpp[0]);
          else if (sx < cn1) term_acc(ar[RX ? sx : 0], sx, F[1], Fp[1], Fpp[1]);
          else term_acc(ar[RX ? sx : 0], sx, F[2], Fp[2], Fpp[2]);
        }
      }
    } else {
      int sx = 0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
#pragma unroll(2)
        for (int mm = 0; mm < CT.cnt[q]; ++mm, ++sx) {
          T a[4];
          ldg_vec<T, 4>(fx + (int64_t)CT.flat[sx] * tstride, a);
          term_acc(a, sx, F[q], Fp[q], Fpp[q]);
        }
      }
    }
  };
  // first output cell of the warp at marching position mm
  auto cell0 = [&](int64_t mm) { return AX == 1 ? i0 + nw0 * (mm + nw1 * row) : i0 + nw0 * (row + nw1 * mm); };
  auto stage_pair = [&](T* arr_a, T* arr_b, int64_t cl0, const T (&A)[3][3], const T (&B)[3][3], T J) {
    bulk_wait_read();  // the TMA unit has finished reading the previous use of the two buffers
    __syncwarp();
    Stage<T, 10> Sa(arr_a, cl0, b10), Sb(arr_b, cl0, b10 + stage_len<T, 10>);
    put10(Sa.slot(lane), A, J);  // every lane stages (no branch); slots >= nvalid are never sent
    put10(Sb.slot(lane), B, J);
    fence_async_smem();
    __syncwarp();
    Sa.issue(nvalid, lane);
    Sb.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  };

  if constexpr (AX == 0) {
    T z[4];
    load_z(m0, z);
    for (int64_t m = m0; m < m1; ++m) {
      __syncwarp();  // every lane has read last step's table
      term_uniform(z);
      __syncwarp();
      load_z(m + 1 < m1 ? m + 1 : m, z);  // next step's factors travel during this step
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Lc[3][3], Rc[3][3], Gf[3][3], Ff[3][3];
      jacobian(F, Fp, Fpp);
      inv3(F, G);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Lc, Rc);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (Lc[a][b] + __shfl_down_sync(FULLMASK, Rc[a][b], 1));
      const T J = rcp(inv3(Gf, Ff));  // det(F_face) = 1 / det(G_face)
      stage_pair(O.face_fwd[0], O.face_inv[0], cell0(m), Ff, Gf, J);
    }
  } else {
    T Lm[3][3];  // L state of the cell of the current step (carried)
    T z[4];
    load_z(m0, z);
    term_uniform(z);
    __syncwarp();
    load_z(m0 + 1, z);  // m0 + 1 <= nw: the factor tables hold nw + 1 entries
    {
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Rm[3][3];
      jacobian(F, Fp, Fpp);
      const T J = inv3(F, G);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Lm, Rm);
      if (AX == 2) stage_pair(O.cell_fwd, O.cell_inv, cell0(m0), F, G, J);
    }
    for (int64_t m = m0; m < m1; ++m) {
      __syncwarp();  // every lane has read the previous table
      term_uniform(z);  // position m + 1
      __syncwarp();
      load_z(m + 2 <= D.nw[MA] ? m + 2 : D.nw[MA], z);
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Ln[3][3], Rn[3][3], Gf[3][3], Ff[3][3];
      jacobian(F, Fp, Fpp);
      const T Jn = inv3(F, G);
      // the cell arrays of the cell m + 1 (the first cell of the next chunk is written by that chunk's prologue)
      if (AX == 2 && m + 1 < m1) stage_pair(O.cell_fwd, O.cell_inv, cell0(m + 1), F, G, Jn);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Ln, Rn);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          Gf[a][b] = T(0.5) * (Lm[a][b] + Rn[a][b]);
          Lm[a][b] = Ln[a][b];
        }
      const T J = rcp(inv3(Gf, Ff));
      stage_pair(O.face_fwd[AX], O.face_inv[AX], cell0(m), Ff, Gf, J);
    }
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// ---------------------------------------------------------------------------
// Round 2, third step (EXPERIMENT, off by default: measured slower, see the launcher): the conserved metrics of face axis
// AX folded into that axis' face kernel (THREE kernels instead of four).  k_mapped_cons_uni is bound by the load / store unit: per pair and plane it fetches six xi values per lane.
// But a lane's xi values never change while its warp marches, and one face axis needs only THREE of the six per pair
// (E0, FD0, E1 for the xi faces; D0, V1, V0 for the eta / zeta faces).  Here they are loaded ONCE per chunk into
// registers (3 x NP doubles, NP <= FOLD_NP pairs), lane p multiplies the warp-uniform row / march products of pair p
// into the five coefficients its face needs, and every lane does five multiply-adds per pair on broadcast reads.  The
// pair loop is fully unrolled (compile-time register indices); the column a pair belongs to is a warp-uniform branch.
// ---------------------------------------------------------------------------
constexpr int FOLD_NP = 12;  // pairs whose xi values fit the registers
constexpr int FOLD_PU = 8;   // coefficients per pair in shared memory (5 + padding to a 16-byte multiple for FP32 too)
#ifndef CG_FOLD_MINB
#define CG_FOLD_MINB 8
#endif
template <class T, int SCH, int AX>
__global__ void __launch_bounds__(32, CG_FOLD_MINB) k_mapped_facecons_uni(Dims D, MCoordTerms CT, MPairs PR, const T* __restrict__ ftab,
                                                                           const T* __restrict__ tab, const T* __restrict__ tabx,
                                                                           int64_t L, int64_t Lx, MetricOut<T> O, int nt, int chunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  constexpr int MA = AX == 1 ? 1 : 2;  // marching axis
  constexpr int RA = AX == 1 ? 2 : 1;  // row axis
  constexpr int OUT = AX == 0 ? 31 : 32;
  const T lam1 = lam1_of<T>(SCH), lam2 = lam2_of<T>(SCH);
  const int64_t L1 = L + 1;
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1];
  const int64_t row = (int64_t)blockIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * OUT;
  const int nvalid = (int)(nw0 - i0 < OUT ? nw0 - i0 : OUT);
  const int64_t i = i0 + lane < nw0 ? i0 + lane : nw0;
  const int64_t m0 = (int64_t)blockIdx.z * chunk;
  const int64_t m1 = m0 + chunk < D.nw[MA] ? m0 + chunk : D.nw[MA];
  const int np = PR.n, cs1 = PR.cstart[1], cs2 = PR.cstart[2];
  T* const b10 = reinterpret_cast<T*>(cg_dyn_smem);     // two Metric staging buffers
  T* const b9 = b10 + 2 * stage_len<T, 10>;              // one ConservedMetric staging buffer
  T* const termU = b9 + stage_len<T, 9>;                 // [nt][UNI_TERM]
  T* const pairU = termU + (int64_t)nt * UNI_TERM;       // [np][FOLD_PU]
  const int64_t tstride = 3 * L1 * 4;
  const T* const fx = ftab + i * 4;
  const int su = lane < nt ? lane : nt - 1;
  const T* const ubase = ftab + (int64_t)CT.flat[su] * tstride;
  T yr[4];
  ldg_vec<T, 4>(ubase + (RA * L1 + row) * 4, yr);
  {
    const T c = T(CT.coef[su]);
#pragma unroll
    for (int o = 0; o < 4; ++o) yr[o] *= c;
  }
  // ---- conserved part: per-lane xi values (registers), per-pair row values of lane p (registers) ----
  const int64_t icx = i0 + lane < nw0 ? i0 + lane : nw0 - 1;
  T xr[FOLD_NP][3];
#pragma unroll
  for (int p = 0; p < FOLD_NP; ++p) {
    const int pp = p < np ? p : np - 1;
    const T* xp = tabx + (int64_t)pp * OP_COUNT * Lx + icx;
    if (AX == 0) { xr[p][0] = __ldg(xp + OP_E0 * Lx); xr[p][1] = __ldg(xp + OP_FD0 * Lx); xr[p][2] = __ldg(xp + OP_E1 * Lx); }
    else { xr[p][0] = __ldg(xp + OP_D0 * Lx); xr[p][1] = __ldg(xp + OP_V1 * Lx); xr[p][2] = __ldg(xp + OP_V0 * Lx); }
  }
  const int pu = lane < np ? lane : np - 1;
  const int64_t pstride = 3 * L * OP_COUNT;
  const T* const prow = tab + (int64_t)pu * pstride;  // + (axis * L + index) * OP_COUNT
  T fr[OP_COUNT];  // operator values of pair `pu` on the ROW axis (fixed), times the pair's coefficient
  ldg_vec<T, OP_COUNT>(prow + (RA * L + row) * OP_COUNT, fr);
  {
    const T C = T(PR.coef[pu]);
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) fr[o] *= C;
  }
  auto load_pz = [&](int64_t mm, T (&z)[OP_COUNT]) { ldg_vec<T, OP_COUNT>(prow + (MA * L + mm) * OP_COUNT, z); };
  // the five coefficients of this face axis (k_mapped_cons_uni: u0..u4 | u5,u6,u9,u10,u13 | u7,u8,u11,u12,u14);
  // y = eta row, z = zeta row of the pair
  auto pair_uniform = [&](const T (&mr)[OP_COUNT]) {
    if (lane < np) {
      const T* const y = AX == 1 ? mr : fr;
      const T* const z = AX == 1 ? fr : mr;
      T u[FOLD_PU];
      if (AX == 0) {
        u[0] = y[OP_V1] * z[OP_D0] - y[OP_D0] * z[OP_V1];
        u[1] = y[OP_V0] * z[OP_V1];
        u[2] = -(z[OP_V0] * y[OP_V1]);
        u[3] = -(y[OP_V0] * z[OP_D0]);
        u[4] = z[OP_V0] * y[OP_D0];
      } else if (AX == 1) {
        u[0] = y[OP_E0] * z[OP_V1];
        u[1] = -(z[OP_V0] * y[OP_E1]);
        u[2] = -(y[OP_E0] * z[OP_D0]);
        u[3] = z[OP_V0] * y[OP_FD0];
        u[4] = y[OP_E1] * z[OP_D0] - y[OP_FD0] * z[OP_V1];
      } else {
        u[0] = -(z[OP_E0] * y[OP_V1]);
        u[1] = y[OP_V0] * z[OP_E1];
        u[2] = z[OP_E0] * y[OP_D0];
        u[3] = -(y[OP_V0] * z[OP_FD0]);
        u[4] = z[OP_FD0] * y[OP_V1] - z[OP_E1] * y[OP_D0];
      }
      u[5] = u[6] = u[7] = T(0);
      sts_vec<T, FOLD_PU>(pairU + lane * FOLD_PU, u);
    }
  };
  // nine conserved entries of this lane's cell from the table in shared memory; Gc[r][c], rows xi / eta / zeta
  auto conserved = [&](T (&Gc)[3][3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) Gc[a][b] = T(0);
#pragma unroll
    for (int p = 0; p < FOLD_NP; ++p) {
      if (p < np) {
        T u[FOLD_PU];
        lds_vec<T, FOLD_PU>(pairU + p * FOLD_PU, u);
        T e0, e1, e2;  // the three rows of the pair's column
        if (AX == 0) {
          e0 = xr[p][0] * u[0];
          e1 = xr[p][1] * u[1] + xr[p][2] * u[3];
          e2 = xr[p][1] * u[2] + xr[p][2] * u[4];
        } else if (AX == 1) {
          e1 = xr[p][0] * u[0] + xr[p][1] * u[2];
          e2 = xr[p][0] * u[1] + xr[p][1] * u[3];
          e0 = xr[p][2] * u[4];
        } else {
          e2 = xr[p][0] * u[0] + xr[p][1] * u[2];
          e1 = xr[p][0] * u[1] + xr[p][1] * u[3];
          e0 = xr[p][2] * u[4];
        }
        if (p < cs1) { Gc[0][0] += e0; Gc[1][0] += e1; Gc[2][0] += e2; }
        else if (p < cs2) { Gc[0][1] += e0; Gc[1][1] += e1; Gc[2][1] += e2; }
        else { Gc[0][2] += e0; Gc[1][2] += e1; Gc[2][2] += e2; }
      }
    }
  };
  auto load_z = [&](int64_t mm, T (&z)[4]) { ldg_vec<T, 4>(ubase + (MA * L1 + mm) * 4, z); };
  auto term_uniform = [&](const T (&z)[4]) {
    if (lane < nt) {
      T u[UNI_TERM];
      u[0] = yr[0] * z[0]; u[1] = yr[1] * z[0]; u[2] = yr[0] * z[1];
      if (AX == 0) {
        u[3] = T(0);
        T v[4] = {u[0], u[1], u[2], u[3]};
        sts_vec<T, 4>(termU + lane * UNI_TERM, v);
      } else {
        u[3] = yr[1] * z[1]; u[4] = yr[0] * z[2]; u[5] = yr[1] * z[2]; u[6] = yr[0] * z[3]; u[7] = T(0);
        sts_vec<T, UNI_TERM>(termU + lane * UNI_TERM, u);
      }
    }
  };
  auto jacobian = [&](T (&F)[3][3], T (&Fp)[3][3], T (&Fpp)[3][3]) {
    int s = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
      for (int d = 0; d < 3; ++d) F[q][d] = Fp[q][d] = Fpp[q][d] = T(0);
#pragma unroll(2)
      for (int mm = 0; mm < CT.cnt[q]; ++mm, ++s) {
        T a[4];
        ldg_vec<T, 4>(fx + (int64_t)CT.flat[s] * tstride, a);
        if (AX == 0) {
          T u[4];
          lds_vec<T, 4>(termU + s * UNI_TERM, u);
          F[q][0] += a[1] * u[0]; F[q][RA] += a[0] * u[1]; F[q][MA] += a[0] * u[2];
          Fp[q][0] += a[2] * u[0]; Fp[q][RA] += a[1] * u[1]; Fp[q][MA] += a[1] * u[2];
          Fpp[q][0] += a[3] * u[0]; Fpp[q][RA] += a[2] * u[1]; Fpp[q][MA] += a[2] * u[2];
        } else {
          T u[UNI_TERM];
          lds_vec<T, UNI_TERM>(termU + s * UNI_TERM, u);
          F[q][0] += a[1] * u[0]; F[q][RA] += a[0] * u[1]; F[q][MA] += a[0] * u[2];
          Fp[q][0] += a[1] * u[2]; Fp[q][RA] += a[0] * u[3]; Fp[q][MA] += a[0] * u[4];
          Fpp[q][0] += a[1] * u[4]; Fpp[q][RA] += a[0] * u[5]; Fpp[q][MA] += a[0] * u[6];
        }
      }
    }
  };
  auto cell0 = [&](int64_t mm) { return AX == 1 ? i0 + nw0 * (mm + nw1 * row) : i0 + nw0 * (row + nw1 * mm); };
  auto stage_pair = [&](T* arr_a, T* arr_b, int64_t cl0, const T (&A)[3][3], const T (&B)[3][3], T J) {
    bulk_wait_read();
    __syncwarp();
    Stage<T, 10> Sa(arr_a, cl0, b10), Sb(arr_b, cl0, b10 + stage_len<T, 10>);
    put10(Sa.slot(lane), A, J);
    put10(Sb.slot(lane), B, J);
    fence_async_smem();
    __syncwarp();
    Sa.issue(nvalid, lane);
    Sb.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  };
  // face arrays + the face's conserved metrics in one group (three TMA stores)
  auto stage_face = [&](int64_t cl0, const T (&Ff)[3][3], const T (&Gf)[3][3], T J, const T (&Gc)[3][3]) {
    bulk_wait_read();
    __syncwarp();
    Stage<T, 10> Sa(O.face_fwd[AX], cl0, b10), Sb(O.face_inv[AX], cl0, b10 + stage_len<T, 10>);
    Stage<T, 9> Sc(O.face_cons[AX], cl0, b9);
    put10(Sa.slot(lane), Ff, J);
    put10(Sb.slot(lane), Gf, J);
    put9(Sc.slot(lane), Gc);
    fence_async_smem();
    __syncwarp();
    Sa.issue(nvalid, lane);
    Sb.issue(nvalid, lane);
    Sc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  };

  T pz[OP_COUNT];  // march-axis operator row of pair `pu` at the position of the NEXT conserved evaluation
  load_pz(m0, pz);
  if constexpr (AX == 0) {
    T z[4];
    load_z(m0, z);
    for (int64_t m = m0; m < m1; ++m) {
      __syncwarp();  // every lane has read last step's tables
      term_uniform(z);
      pair_uniform(pz);
      __syncwarp();
      load_z(m + 1 < m1 ? m + 1 : m, z);
      load_pz(m + 1 < m1 ? m + 1 : m, pz);
      T Gc[3][3];
      conserved(Gc);
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Lc[3][3], Rc[3][3], Gf[3][3], Ff[3][3];
      jacobian(F, Fp, Fpp);
      inv3(F, G);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Lc, Rc);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) Gf[a][b] = T(0.5) * (Lc[a][b] + __shfl_down_sync(FULLMASK, Rc[a][b], 1));
      const T J = rcp(inv3(Gf, Ff));
      stage_face(cell0(m), Ff, Gf, J, Gc);
    }
  } else {
    T Lm[3][3];
    T z[4];
    load_z(m0, z);
    term_uniform(z);
    __syncwarp();
    load_z(m0 + 1, z);
    {
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Rm[3][3];
      jacobian(F, Fp, Fpp);
      const T J = inv3(F, G);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Lm, Rm);
      if (AX == 2) stage_pair(O.cell_fwd, O.cell_inv, cell0(m0), F, G, J);
    }
    for (int64_t m = m0; m < m1; ++m) {
      __syncwarp();
      term_uniform(z);   // Jacobian tables: position m + 1
      pair_uniform(pz);  // conserved tables: position m
      __syncwarp();
      load_z(m + 2 <= D.nw[MA] ? m + 2 : D.nw[MA], z);
      load_pz(m + 1 < m1 ? m + 1 : m, pz);
      T Gc[3][3];
      conserved(Gc);
      T F[3][3], Fp[3][3], Fpp[3][3], G[3][3], Ln[3][3], Rn[3][3], Gf[3][3], Ff[3][3];
      jacobian(F, Fp, Fpp);
      const T Jn = inv3(F, G);
      if (AX == 2 && m + 1 < m1) stage_pair(O.cell_fwd, O.cell_inv, cell0(m + 1), F, G, Jn);
      mapped_states3_g(G, Fp, Fpp, lam1, lam2, Ln, Rn);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          Gf[a][b] = T(0.5) * (Lm[a][b] + Rn[a][b]);
          Lm[a][b] = Ln[a][b];
        }
      const T J = rcp(inv3(Gf, Ff));
      stage_face(cell0(m), Ff, Gf, J, Gc);
    }
  }
  if (lane == 0) bulk_wait_read();
}

// ---------------------------------------------------------------------------
// Conserved metrics of ONE face axis (second session of round 2): k_mapped_cons_uni split by face axis.  A face axis needs
// only three of a pair's six xi values (E0, FD0, E1 for the xi faces; D0, V1, V0 for the eta / zeta faces) and five of the
// fifteen warp-uniform coefficients, and a lane's xi values never change while its warp marches along k: they are loaded
// ONCE per chunk into registers (3 x FOLD_NP doubles), so the loop has no table loads at all — per pair three broadcast
// LDS.128 and five multiply-adds.  ~100 registers, 16 one-warp blocks per SM; LSU wavefronts per cell and plane over the
// three launches 162 against 222 of k_mapped_cons_uni.  (Folding the same loop into the face kernels lost, see
// k_mapped_facecons_uni: there the 36 doubles came on top of the Jacobian states.)
// ---------------------------------------------------------------------------
#ifndef CG_CONSF_MINB
#define CG_CONSF_MINB 16
#endif
template <class T, int AX>
__global__ void __launch_bounds__(32, CG_CONSF_MINB) k_mapped_cons_face(Dims D, MPairs PR, const T* __restrict__ tab,
                                                                         const T* __restrict__ tabx, int64_t L, int64_t Lx,
                                                                         MetricOut<T> O, int kchunk) {
  extern __shared__ __align__(128) unsigned char cg_dyn_smem[];
  const int lane = threadIdx.x;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const int64_t j = (int64_t)blockIdx.x;
  const int64_t i0 = (int64_t)blockIdx.y * 32;
  const int nvalid = (int)(nw0 - i0 < 32 ? nw0 - i0 : 32);
  const int64_t ic = i0 + lane < nw0 ? i0 + lane : nw0 - 1;  // lanes past the row end repeat the last cell (never sent)
  const int64_t k0 = (int64_t)blockIdx.z * kchunk;
  const int64_t k1 = k0 + kchunk < nw2 ? k0 + kchunk : nw2;
  constexpr int PU = sizeof(T) == 8 ? 6 : 8;  // coefficients per pair in shared memory (5 + padding to 16-byte pieces)
  const int np = PR.n, cs1 = PR.cstart[1], cs2 = PR.cstart[2];
  T* const b9 = reinterpret_cast<T*>(cg_dyn_smem);  // TWO ConservedMetric staging buffers (a step never waits for the
  T* const pairU = b9 + 2 * stage_len<T, 9>;         // TMA read of the step before) | [np][PU]
  T xr[FOLD_NP][3];
#pragma unroll
  for (int p = 0; p < FOLD_NP; ++p) {
    const int pp = p < np ? p : np - 1;
    const T* xp = tabx + (int64_t)pp * OP_COUNT * Lx + ic;
    if (AX == 0) { xr[p][0] = __ldg(xp + OP_E0 * Lx); xr[p][1] = __ldg(xp + OP_FD0 * Lx); xr[p][2] = __ldg(xp + OP_E1 * Lx); }
    else { xr[p][0] = __ldg(xp + OP_D0 * Lx); xr[p][1] = __ldg(xp + OP_V1 * Lx); xr[p][2] = __ldg(xp + OP_V0 * Lx); }
  }
  const int pu = lane < np ? lane : np - 1;
  const T* const prow = tab + (int64_t)pu * 3 * L * OP_COUNT;
  T y[OP_COUNT];  // eta row of pair `pu` (fixed for the block), times the pair's coefficient
  ldg_vec<T, OP_COUNT>(prow + (L + j) * OP_COUNT, y);
  {
    const T C = T(PR.coef[pu]);
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) y[o] *= C;
  }
  T z[OP_COUNT];
  ldg_vec<T, OP_COUNT>(prow + (2 * L + k0) * OP_COUNT, z);
  for (int64_t m = k0; m < k1; ++m) {
    __syncwarp();  // every lane has read last step's coefficients
    if (lane < np) {
      T u[PU];
      if (AX == 0) {
        u[0] = y[OP_V1] * z[OP_D0] - y[OP_D0] * z[OP_V1];
        u[1] = y[OP_V0] * z[OP_V1];
        u[2] = -(z[OP_V0] * y[OP_V1]);
        u[3] = -(y[OP_V0] * z[OP_D0]);
        u[4] = z[OP_V0] * y[OP_D0];
      } else if (AX == 1) {
        u[0] = y[OP_E0] * z[OP_V1];
        u[1] = -(z[OP_V0] * y[OP_E1]);
        u[2] = -(y[OP_E0] * z[OP_D0]);
        u[3] = z[OP_V0] * y[OP_FD0];
        u[4] = y[OP_E1] * z[OP_D0] - y[OP_FD0] * z[OP_V1];
      } else {
        u[0] = -(z[OP_E0] * y[OP_V1]);
        u[1] = y[OP_V0] * z[OP_E1];
        u[2] = z[OP_E0] * y[OP_D0];
        u[3] = -(y[OP_V0] * z[OP_FD0]);
        u[4] = z[OP_FD0] * y[OP_V1] - z[OP_E1] * y[OP_D0];
      }
#pragma unroll
      for (int o = 5; o < PU; ++o) u[o] = T(0);
      sts_vec<T, PU>(pairU + lane * PU, u);
    }
    __syncwarp();
    ldg_vec<T, OP_COUNT>(prow + (2 * L + (m + 1 < k1 ? m + 1 : m)) * OP_COUNT, z);  // next step's zeta row travels now
    T Gc[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) Gc[a][b] = T(0);
#pragma unroll
    for (int p = 0; p < FOLD_NP; ++p) {
      if (p < np) {
        T u[PU];
        lds_vec<T, PU>(pairU + p * PU, u);
        // the pair's column is warp-uniform; multiply-adds straight into that column's three rows (xi / eta / zeta)
        auto acc = [&](T& g0, T& g1, T& g2) {
          if (AX == 0) {
            g0 += xr[p][0] * u[0];
            g1 += xr[p][1] * u[1] + xr[p][2] * u[3];
            g2 += xr[p][1] * u[2] + xr[p][2] * u[4];
          } else if (AX == 1) {
            g1 += xr[p][0] * u[0] + xr[p][1] * u[2];
            g2 += xr[p][0] * u[1] + xr[p][1] * u[3];
            g0 += xr[p][2] * u[4];
          } else {
            g2 += xr[p][0] * u[0] + xr[p][1] * u[2];
            g1 += xr[p][0] * u[1] + xr[p][1] * u[3];
            g0 += xr[p][2] * u[4];
          }
        };
        if (p < cs1) acc(Gc[0][0], Gc[1][0], Gc[2][0]);
        else if (p < cs2) acc(Gc[0][1], Gc[1][1], Gc[2][1]);
        else acc(Gc[0][2], Gc[1][2], Gc[2][2]);
      }
    }
    bulk_wait_read1();  // the TMA unit has finished reading the buffer of TWO steps ago (lanes > 0 own no groups)
    __syncwarp();
    Stage<T, 9> Sc(O.face_cons[AX], i0 + nw0 * (j + nw1 * m), b9 + (int)((m - k0) & 1) * stage_len<T, 9>);
    put9(Sc.slot(lane), Gc);  // every lane stages (no branch); slots >= nvalid are never sent
    fence_async_smem();
    __syncwarp();
    Sc.issue(nvalid, lane);
    if (lane == 0) bulk_commit();
  }
  if (lane == 0) bulk_wait_read();  // shared memory must outlive the TMA reads
}

// launches the four kernels on `st`; returns the number of launches, 0 when the configuration is not covered
template <class T>
inline int launch_mapped_metrics3d_split(const Dims& D, const MTerms& TM, const MPairs& PR, const T* tab, const T* tabx,
                                         const T* ftab, int64_t L, const MetricOut<T>& O, int scheme, cudaStream_t st) {
  const int nt = TM.n, np = PR.n;
  if (nt < 1 || nt > 32 || np < 1 || !tabx) return 0;
  const int64_t nw0 = D.nw[0], nw1 = D.nw[1], nw2 = D.nw[2];
  const unsigned g32 = (unsigned)((nw0 + 31) / 32), g31 = (unsigned)((nw0 + 30) / 31);
  if (g31 > 65535u) return 0;
  const MCoordTerms CT = coord_terms(TM);
  bool ok = true;
  // experiment switch CG_MAPPED_FOLD=1 (listed in include/curvigrid_cuda.h): the conserved metrics folded into the three
  // face kernels.  Correct (the MappedGrid tests pass on it) but SLOWER than the four kernels: 212-254 registers leave 8
  // warps per SM and each face kernel grows by 3.6-5.0 ms (9.3 + 9.1 + 10.7 = 29.1 ms against 9.9 + 4.9 + 4.1 + 7.1 =
  // 26.0 ms at 512^3), more than its 3.3 ms share of k_mapped_cons_uni.  OFF by default.
  static const bool fold_env = getenv("CG_MAPPED_FOLD") && atoi(getenv("CG_MAPPED_FOLD")) == 1;
  if (fold_env && np <= FOLD_NP && PR.cstart[1] <= PR.cstart[2]) {
    const int fsm = (int)((2 * stage_len<T, 10> + stage_len<T, 9> + (int64_t)nt * UNI_TERM + (int64_t)np * FOLD_PU) * sizeof(T));
    auto fface = [&](auto kern, unsigned gx, int64_t rows, int64_t extent) {
      const int chunk = pick_chunk((int64_t)gx * rows, extent, 148LL * CG_FOLD_MINB, 32, 1.0);
      ok &= cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm) == cudaSuccess;
      kern<<<dim3((unsigned)rows, gx, (unsigned)((extent + chunk - 1) / chunk)), 32, fsm, st>>>(D, CT, PR, ftab, tab, tabx, L, mapped_lx(L), O, nt, chunk);
    };
    auto fall = [&](auto sch) {
      constexpr int S = decltype(sch)::value;
      fface(k_mapped_facecons_uni<T, S, 0>, g31, nw1, nw2);
      fface(k_mapped_facecons_uni<T, S, 1>, g32, nw2, nw1);
      fface(k_mapped_facecons_uni<T, S, 2>, g32, nw1, nw2);
    };
    if (scheme == ENDPOINT) fall(std::integral_constant<int, ENDPOINT>{});
    else if (scheme == GRADIENT) fall(std::integral_constant<int, GRADIENT>{});
    else fall(std::integral_constant<int, CURVATURE>{});
    return ok ? 3 : -1;
  }
  // debug switch CG_MAPPED_CONS=uni (listed in include/curvigrid_cuda.h): the single conserved kernel of all three face axes
  static const bool cons_uni_env = getenv("CG_MAPPED_CONS") && std::string(getenv("CG_MAPPED_CONS")) == "uni";
  if (!cons_uni_env && np <= FOLD_NP) {
    const int smem = (int)((2 * stage_len<T, 9> + (int64_t)np * (sizeof(T) == 8 ? 6 : 8)) * sizeof(T));
    const int kchunk = pick_chunk((int64_t)g32 * nw1, nw2, 148LL * CG_CONSF_MINB, 32, 0.5);
    const dim3 grid((unsigned)nw1, g32, (unsigned)((nw2 + kchunk - 1) / kchunk));
    auto cons = [&](auto kern) {
      ok &= cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
      kern<<<grid, 32, smem, st>>>(D, PR, tab, tabx, L, mapped_lx(L), O, kchunk);
    };
    cons(k_mapped_cons_face<T, 0>);
    cons(k_mapped_cons_face<T, 1>);
    cons(k_mapped_cons_face<T, 2>);
  } else {
    const int smem = (int)(CONS_KS * (3 * stage_len<T, 9> + (int64_t)np * UNI_PAIR) * sizeof(T));
    int per_sm = (int)(200 * 1024 / (smem + 1024));
    if (per_sm > CG_SPLIT_CONS_MINB) per_sm = CG_SPLIT_CONS_MINB;
    if (per_sm < 1) return 0;
    int kchunk = pick_chunk((int64_t)g32 * nw1, nw2, 148LL * per_sm, 32, 0.5);
    kchunk = (kchunk + CONS_KS - 1) / CONS_KS * CONS_KS;  // whole passes
    ok &= cudaFuncSetAttribute(k_mapped_cons_uni<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess;
    k_mapped_cons_uni<T><<<dim3((unsigned)nw1, g32, (unsigned)((nw2 + kchunk - 1) / kchunk)), 32, smem, st>>>(D, PR, tab, tabx, L, mapped_lx(L), O, kchunk);
  }
  const int fsmem = (int)((2 * stage_len<T, 10> + (int64_t)nt * UNI_TERM) * sizeof(T));
  auto face = [&](auto kern, unsigned gx, int64_t rows, int64_t extent) {
    const int chunk = pick_chunk((int64_t)gx * rows, extent, 148LL * CG_SPLIT_FACE_MINB, 32, 1.0);
    ok &= cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem) == cudaSuccess;
    kern<<<dim3((unsigned)rows, gx, (unsigned)((extent + chunk - 1) / chunk)), 32, fsmem, st>>>(D, CT, ftab, L, O, nt, chunk);
  };
  // debug switch CG_MAPPED_FACE_RX=0 (listed in include/curvigrid_cuda.h): xi factor rows re-read from L1 in every step
  static const bool rx_env = !(getenv("CG_MAPPED_FACE_RX") && atoi(getenv("CG_MAPPED_FACE_RX")) == 0);
  auto all = [&](auto sch) {
    constexpr int S = decltype(sch)::value;
    if (rx_env && nt <= FACE_NT) {
      face(k_mapped_face_uni<T, S, 0, true>, g31, nw1, nw2);
      face(k_mapped_face_uni<T, S, 1, true>, g32, nw2, nw1);
      face(k_mapped_face_uni<T, S, 2, true>, g32, nw1, nw2);
    } else {
      face(k_mapped_face_uni<T, S, 0, false>, g31, nw1, nw2);
      face(k_mapped_face_uni<T, S, 1, false>, g32, nw2, nw1);
      face(k_mapped_face_uni<T, S, 2, false>, g32, nw1, nw2);
    }
  };
  if (scheme == ENDPOINT) all(std::integral_constant<int, ENDPOINT>{});
  else if (scheme == GRADIENT) all(std::integral_constant<int, GRADIENT>{});
  else all(std::integral_constant<int, CURVATURE>{});
  return ok ? (!cons_uni_env && np <= FOLD_NP ? 6 : 4) : -1;
}

}  // namespace cg
