// FFMA engine of the fused step kernel: the first correct path (round 1, v1/v2 in DESIGN.md section 3), kept as the A/B
// baseline that justifies running the skinny mat-vecs on the tensor cores.  NOT part of the product library: the file
// is empty unless compiled with -DPINN_AB_BUILD (tools/build_ab.sh builds a separate libpinn_b200_ab.so for
// bench.py --library ...); the product has one engine and no run-time dispatch.
//
// What is computed (closed form of the reference's autograd path, oracle/closed_form.py):
//   psi, lap psi, residual, loss sums and dLtot/dtheta of the parametric H2+ model
//   (poc/main.py:247-303, 82-120, 341-355 and train.py:41-57).
//
// Work decomposition
//   * a GROUP of warps owns a tile of 32 collocation points (lane = point).  The warps of a
//     group are specialised by role: one warp per MLP evaluation of the base network
//     (poc: (f1,f2) and the inversion image (f2,f1), poc/main.py:255-260; train.py: one)
//     and one warp for the E(R) network + gate.  Roles meet once per tile through a
//     shared-memory mailbox and a named barrier (forward results -> reverse-sweep seeds).
//   * skinny layers (16x16 x 4 Taylor channels, 32x32) run as FFMA mat-vecs, weights read as
//     broadcast LDS.128 from the shared-memory weight image that is staged ONCE per CTA
//     with a TMA bulk copy (cp.async.bulk + mbarrier).
//   * the weight-gradient contractions dW2 = sum_p G_p^T H_p and dWE2 = sum_p V_p^T E1_p
//     have K = points and are dense: they run on the tensor cores (mma.sync m16n8k8
//     TF32 with the 3xTF32 split, fp32-accurate), operands staged per warp in shared memory,
//     accumulators persistent in registers over all tiles of the launch.
//   * bias/vector gradients and the loss sums are reduced over the 32 points of a tile by column sums over
//     the swizzled shared-memory stash (colsum).
//   * per-CTA results go to a partial row in global memory (pinn_reduce.cu adds the rows).
#ifdef PINN_AB_BUILD
#include "pinn_device.cuh"
#include "pinn_sample.cuh"
#include "pinn_train.h"

namespace pinn {


// ---------------------------------------------------------------------------------------------
// base MLP, one evaluation at (a,b): 4-channel Taylor forward (value, d/da, d/db, D)
//   Hs row (this lane's point): h_c[k]  at column c*16+k   (kept for the dW2 contraction and the reverse sweep)
//   Gs row:                     t_j, va_j, vb_j, vD_j at column c*16+j  (overwritten by the adjoints later)
// ---------------------------------------------------------------------------------------------
template <bool STASH>
__device__ __forceinline__ void mlp_forward(const Wts& w, float a, float b, float al1, float al2, float al11,
                                            float al12, float al22, float* __restrict__ Hrow,
                                            float* __restrict__ Grow, int sx, float& Nv, float& Dv) {
  float h[4][NH];
#pragma unroll
  for (int k4 = 0; k4 < NH; k4 += 4) {
    const float4 w0v = LD4(&w.w0[k4]), w1v = LD4(&w.w1[k4]), b1v = LD4(&w.b1[k4]);
    const float4 q00 = LD4(&w.ww00[k4]), q01 = LD4(&w.ww01[k4]), q11 = LD4(&w.ww11[k4]);
    const float w0a[4] = {w0v.x, w0v.y, w0v.z, w0v.w}, w1a[4] = {w1v.x, w1v.y, w1v.z, w1v.w};
    const float b1a[4] = {b1v.x, b1v.y, b1v.z, b1v.w};
    const float q00a[4] = {q00.x, q00.y, q00.z, q00.w}, q01a[4] = {q01.x, q01.y, q01.z, q01.w};
    const float q11a[4] = {q11.x, q11.y, q11.z, q11.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = k4 + i;
      const float u = fmaf(a, w0a[i], fmaf(b, w1a[i], b1a[i]));
      const float s = sigm(u);
      const float sp = fmaf(-s, s, s);             // s(1-s)
      const float spp = fmaf(-2.0f * s, sp, sp);   // s'(1-2s)
      const float d1 = fmaf(al1, w0a[i], al2 * w1a[i]);
      const float q = fmaf(al11, q00a[i], fmaf(al12, q01a[i], al22 * q11a[i]));
      h[0][k] = s;
      h[1][k] = sp * w0a[i];
      h[2][k] = sp * w1a[i];
      h[3][k] = fmaf(sp, d1, spp * q);
    }
  }
  if (STASH) {
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int k4 = 0; k4 < NH; k4 += 4)
        ST4(&Hrow[(c * NH + k4) ^ sx], h[c][k4], h[c][k4 + 1], h[c][k4 + 2], h[c][k4 + 3]);
  }
  float accN = 0.0f, accD = 0.0f;
#pragma unroll 1
  for (int j4 = 0; j4 < NH; j4 += 4) {
    float tt[4], va[4], vb[4], vD[4];
    const float4 b2v = LD4(&w.b2[j4]), wov = LD4(&w.wo[j4]);
    const float b2a[4] = {b2v.x, b2v.y, b2v.z, b2v.w}, woa[4] = {wov.x, wov.y, wov.z, wov.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float* wrow = &w.W2[(j4 + i) * NH];
      float v0 = b2a[i], v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < NH; k4 += 4) {
        const float4 wv = LD4(&wrow[k4]);
        const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          v0 = fmaf(wa[kk], h[0][k4 + kk], v0);
          v1 = fmaf(wa[kk], h[1][k4 + kk], v1);
          v2 = fmaf(wa[kk], h[2][k4 + kk], v2);
          v3 = fmaf(wa[kk], h[3][k4 + kk], v3);
        }
      }
      const float t = sigm(v0);
      const float tp = fmaf(-t, t, t);
      const float tpp = fmaf(-2.0f * t, tp, tp);
      const float Q = fmaf(al11 * v1, v1, fmaf(al12 * v1, v2, al22 * v2 * v2));
      const float gD = fmaf(tp, v3, tpp * Q);
      accN = fmaf(woa[i], t, accN);
      accD = fmaf(woa[i], gD, accD);
      tt[i] = t; va[i] = v1; vb[i] = v2; vD[i] = v3;
    }
    if (STASH) {
      ST4(&Grow[(0 * NH + j4) ^ sx], tt[0], tt[1], tt[2], tt[3]);
      ST4(&Grow[(1 * NH + j4) ^ sx], va[0], va[1], va[2], va[3]);
      ST4(&Grow[(2 * NH + j4) ^ sx], vb[0], vb[1], vb[2], vb[3]);
      ST4(&Grow[(3 * NH + j4) ^ sx], vD[0], vD[1], vD[2], vD[3]);
    }
  }
  Nv = accN;
  Dv = accD;
}

// persistent per-lane accumulators of an MLP warp
struct MlpAcc {
  float cW2[2][4];   // mma C fragments of dW2 (16x16): n-tile 0/1
  float s0;          // lane<16: db2[lane]        lane>=16: dwo[lane-16]
  float s1;          // lane<16: dW1[lane][0]     lane>=16: dW1[lane-16][1]
  float s2;          // lane<16: db1[lane]        lane>=16: extra[lane-16] (loss sums, role 0)
};

// Reverse sweep of one MLP evaluation for the seeds lamN = dL/dNv, lamD = dL/dDv
// (oracle/closed_form.py:mlp_bwd).  `extra` are 8 more per-point values summed over the tile.
__device__ __forceinline__ void mlp_backward(const Wts& w, float a, float b, float al1, float al2, float al11,
                                             float al12, float al22, float lamN, float lamD,
                                             float* __restrict__ Hs, float* __restrict__ Gs, int lane,
                                             const float (&extra)[8], MlpAcc& acc) {
  const int sx = swz(lane);
  float* Hrow = Hs + lane * ROWH;
  float* Grow = Gs + lane * ROWH;
  float vbar[4][NH];
  float dwo[NH];
#pragma unroll
  for (int j4 = 0; j4 < NH; j4 += 4) {
    const float4 tv = LD4(&Grow[(0 * NH + j4) ^ sx]), av = LD4(&Grow[(1 * NH + j4) ^ sx]);
    const float4 bv = LD4(&Grow[(2 * NH + j4) ^ sx]), dv = LD4(&Grow[(3 * NH + j4) ^ sx]);
    const float4 wov = LD4(&w.wo[j4]);
    const float ta[4] = {tv.x, tv.y, tv.z, tv.w}, vaa[4] = {av.x, av.y, av.z, av.w};
    const float vba[4] = {bv.x, bv.y, bv.z, bv.w}, vDa[4] = {dv.x, dv.y, dv.z, dv.w};
    const float woa[4] = {wov.x, wov.y, wov.z, wov.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int j = j4 + i;
      const float t = ta[i], v1 = vaa[i], v2 = vba[i], v3 = vDa[i];
      const float tp = fmaf(-t, t, t);
      const float tpp = fmaf(-2.0f * t, tp, tp);
      const float tppp = tp * fmaf(-6.0f, tp, 1.0f);
      const float Q = fmaf(al11 * v1, v1, fmaf(al12 * v1, v2, al22 * v2 * v2));
      const float gD = fmaf(tp, v3, tpp * Q);
      const float tbar = lamN * woa[i], gDbar = lamD * woa[i];
      dwo[j] = fmaf(lamN, t, lamD * gD);
      const float c2 = gDbar * tpp;
      vbar[3][j] = gDbar * tp;
      vbar[1][j] = c2 * fmaf(2.0f * al11, v1, al12 * v2);
      vbar[2][j] = c2 * fmaf(al12, v1, 2.0f * al22 * v2);
      vbar[0][j] = fmaf(tbar, tp, gDbar * fmaf(tpp, v3, tppp * Q));
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
      ST4(&Grow[(c * NH + j4) ^ sx], vbar[c][j4], vbar[c][j4 + 1], vbar[c][j4 + 2], vbar[c][j4 + 3]);
  }
  __syncwarp();

  // ---- dW2[j][k] += sum_{p,c} G[p][c][j] * H[p][c][k] on the tensor cores (3xTF32) ----
  {
    const int g = lane >> 2, t = lane & 3, tx = t << 3;
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
#pragma unroll
      for (int po = 0; po < 4; po++) {
        const float* Ga = Gs + (po * 8 + t) * ROWH;
        const float* Gb = Ga + 4 * ROWH;
        const float* Ha = Hs + (po * 8 + t) * ROWH;
        const float* Hb = Ha + 4 * ROWH;
        const int ca = (c * NH + g) ^ tx, cb = (c * NH + g + 8) ^ tx;
        uint32_t ah[4], al[4];
        split_tf32(Ga[ca], ah[0], al[0]);
        split_tf32(Ga[cb], ah[1], al[1]);
        split_tf32(Gb[ca], ah[2], al[2]);
        split_tf32(Gb[cb], ah[3], al[3]);
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(Ha[ca], bh0, bl0);
        split_tf32(Hb[ca], bh1, bl1);
        mma_3xtf32(acc.cW2[0], ah, al, bh0, bh1, bl0, bl1);
        split_tf32(Ha[cb], bh0, bl0);
        split_tf32(Hb[cb], bh1, bl1);
        mma_3xtf32(acc.cW2[1], ah, al, bh0, bh1, bl0, bl1);
      }
    }
  }
  __syncwarp();  // all fragment loads done: channel 1..3 regions of the stash rows may be reused
  // ---- db2 (= column sums of vbar channel 0) and dwo ----
#pragma unroll
  for (int j4 = 0; j4 < NH; j4 += 4) ST4(&Grow[(NH + j4) ^ sx], dwo[j4], dwo[j4 + 1], dwo[j4 + 2], dwo[j4 + 3]);
  __syncwarp();
  acc.s0 += colsum<ROWH>(Gs, lane);

  // ---- hbar = W2^T vbar, then layer-1 reverse sweep, 4 neurons per iteration ----
#pragma unroll 1
  for (int k4 = 0; k4 < NH; k4 += 4) {
    const float4 sv = LD4(&Hrow[k4 ^ sx]);  // h channel 0 = s_k
    const float4 w0v = LD4(&w.w0[k4]), w1v = LD4(&w.w1[k4]);
    const float4 q00 = LD4(&w.ww00[k4]), q01 = LD4(&w.ww01[k4]), q11 = LD4(&w.ww11[k4]);
    const float sa[4] = {sv.x, sv.y, sv.z, sv.w};
    const float w0a[4] = {w0v.x, w0v.y, w0v.z, w0v.w}, w1a[4] = {w1v.x, w1v.y, w1v.z, w1v.w};
    const float q00a[4] = {q00.x, q00.y, q00.z, q00.w}, q01a[4] = {q01.x, q01.y, q01.z, q01.w};
    const float q11a[4] = {q11.x, q11.y, q11.z, q11.w};
    float o0[4], o1[4], ob[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float* wrow = &w.W2T[(k4 + i) * NH];
      float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f, h3 = 0.0f;
#pragma unroll
      for (int j4 = 0; j4 < NH; j4 += 4) {
        const float4 wv = LD4(&wrow[j4]);
        const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          h0 = fmaf(wa[jj], vbar[0][j4 + jj], h0);
          h1 = fmaf(wa[jj], vbar[1][j4 + jj], h1);
          h2 = fmaf(wa[jj], vbar[2][j4 + jj], h2);
          h3 = fmaf(wa[jj], vbar[3][j4 + jj], h3);
        }
      }
      const float s = sa[i], w0 = w0a[i], w1 = w1a[i];
      const float sp = fmaf(-s, s, s);
      const float spp = fmaf(-2.0f * s, sp, sp);
      const float sppp = sp * fmaf(-6.0f, sp, 1.0f);
      const float d1 = fmaf(al1, w0, al2 * w1);
      const float q = fmaf(al11, q00a[i], fmaf(al12, q01a[i], al22 * q11a[i]));
      const float ubar = fmaf(h0, sp, fmaf(fmaf(h1, w0, fmaf(h2, w1, h3 * d1)), spp, h3 * q * sppp));
      o0[i] = fmaf(h1, sp, fmaf(h3, fmaf(sp, al1, spp * fmaf(2.0f * al11, w0, al12 * w1)), ubar * a));
      o1[i] = fmaf(h2, sp, fmaf(h3, fmaf(sp, al2, spp * fmaf(al12, w0, 2.0f * al22 * w1)), ubar * b));
      ob[i] = ubar;
    }
    ST4(&Hrow[(1 * NH + k4) ^ sx], o0[0], o0[1], o0[2], o0[3]);
    ST4(&Hrow[(2 * NH + k4) ^ sx], o1[0], o1[1], o1[2], o1[3]);
    ST4(&Hrow[(3 * NH + k4) ^ sx], ob[0], ob[1], ob[2], ob[3]);
  }
  ST4(&Hrow[0 ^ sx], extra[0], extra[1], extra[2], extra[3]);
  ST4(&Hrow[4 ^ sx], extra[4], extra[5], extra[6], extra[7]);
  __syncwarp();
  acc.s1 += colsum<ROWH>(Hs, NH + lane);                          // columns 16..47: dw0 | dw1
  acc.s2 += colsum<ROWH>(Hs, lane < 16 ? 3 * NH + lane : lane - 16);  // columns 48..63: db1 ; 0..15: extras
  __syncwarp();
}

// loss sums only (fine-tune mode: no base-MLP reverse sweep); role 0
__device__ __forceinline__ void mlp_extras_only(float* __restrict__ Hs, int lane, const float (&extra)[8], MlpAcc& acc) {
  const int sx = swz(lane);
  float* Hrow = Hs + lane * ROWH;
  __syncwarp();
  ST4(&Hrow[0 ^ sx], extra[0], extra[1], extra[2], extra[3]);
  ST4(&Hrow[4 ^ sx], extra[4], extra[5], extra[6], extra[7]);
  __syncwarp();
  const float v = colsum<ROWH>(Hs, lane & 7);
  if (lane >= 16 && lane < 24) acc.s2 += v;
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// E(R) network (poc/main.py:249-253; train.py:50-52) and gate (poc/main.py:262-264; train.py:48-49)
// ---------------------------------------------------------------------------------------------
template <bool STASH>
__device__ __forceinline__ float enet_forward(const Wts& w, float R, float* __restrict__ E1row,
                                              float* __restrict__ Vrow, int sx) {
  float e1[NE];
#pragma unroll
  for (int k4 = 0; k4 < NE; k4 += 4) {
    const float4 wv = LD4(&w.WE1[k4]), bv = LD4(&w.bE1[k4]);
    e1[k4 + 0] = sigm(fmaf(R, wv.x, bv.x));
    e1[k4 + 1] = sigm(fmaf(R, wv.y, bv.y));
    e1[k4 + 2] = sigm(fmaf(R, wv.z, bv.z));
    e1[k4 + 3] = sigm(fmaf(R, wv.w, bv.w));
    if (STASH) ST4(&E1row[k4 ^ sx], e1[k4], e1[k4 + 1], e1[k4 + 2], e1[k4 + 3]);
  }
  float E = w.bE;
#pragma unroll 1
  for (int j4 = 0; j4 < NE; j4 += 4) {
    const float4 b2v = LD4(&w.bE2[j4]), wEv = LD4(&w.wE[j4]);
    const float b2a[4] = {b2v.x, b2v.y, b2v.z, b2v.w}, wEa[4] = {wEv.x, wEv.y, wEv.z, wEv.w};
    float e2[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float* wrow = &w.WE2[(j4 + i) * NE];
      float v0 = b2a[i], v1 = 0.0f;  // two partial sums for ILP
#pragma unroll
      for (int k4 = 0; k4 < NE; k4 += 8) {
        const float4 wa = LD4(&wrow[k4]), wb = LD4(&wrow[k4 + 4]);
        v0 = fmaf(wa.x, e1[k4 + 0], v0); v0 = fmaf(wa.y, e1[k4 + 1], v0);
        v0 = fmaf(wa.z, e1[k4 + 2], v0); v0 = fmaf(wa.w, e1[k4 + 3], v0);
        v1 = fmaf(wb.x, e1[k4 + 4], v1); v1 = fmaf(wb.y, e1[k4 + 5], v1);
        v1 = fmaf(wb.z, e1[k4 + 6], v1); v1 = fmaf(wb.w, e1[k4 + 7], v1);
      }
      e2[i] = sigm(v0 + v1);
      E = fmaf(wEa[i], e2[i], E);
    }
    if (STASH) ST4(&Vrow[j4 ^ sx], e2[0], e2[1], e2[2], e2[3]);
  }
  return E;
}

__device__ __forceinline__ float gate_forward(const Wts& w, float R) {
  float g = w.bg;
#pragma unroll
  for (int i = 0; i < NL; i++) g = fmaf(w.wg[i], sigm(fmaf(R, w.WgL[i], w.bgL[i])), g);
  return g;
}

struct EnetAcc {
  float cWE2[2][4][4];         // mma C fragments of dWE2 (32x32): [m-tile][n-tile]
  float s0, s1, s2, s3, s4;    // per lane: dwE[lane], dbE2[lane], dWE1[lane], dbE1[lane], {dWgL,dbgL,dwg,dbg,dbE}[lane]
};

// Reverse sweep of the E-net and gate for the seeds Ebar = dL/dE, gbar = dL/dgate.
__device__ __forceinline__ void enet_backward(const Wts& w, float R, float Ebar, float gbar, bool gate_grads,
                                              float* __restrict__ E1s, float* __restrict__ Vs, int lane,
                                              EnetAcc& acc) {
  const int sx = swz(lane);
  float* E1row = E1s + lane * ROWE;
  float* Vrow = Vs + lane * ROWE;
  __syncwarp();
  // ---- dwE[j] = sum_p Ebar_p e2[p][j] (Vs still holds e2) ----
  {
    float plain, weighted;
    colsum_w<ROWE>(Vs, lane, Ebar, plain, weighted);
    acc.s0 += weighted;
  }
  __syncwarp();
  float vbar[NE];
#pragma unroll
  for (int j4 = 0; j4 < NE; j4 += 4) {
    const float4 ev = LD4(&Vrow[j4 ^ sx]), wEv = LD4(&w.wE[j4]);
    const float ea[4] = {ev.x, ev.y, ev.z, ev.w}, wEa[4] = {wEv.x, wEv.y, wEv.z, wEv.w};
#pragma unroll
    for (int i = 0; i < 4; i++) vbar[j4 + i] = Ebar * wEa[i] * fmaf(-ea[i], ea[i], ea[i]);
    ST4(&Vrow[j4 ^ sx], vbar[j4], vbar[j4 + 1], vbar[j4 + 2], vbar[j4 + 3]);
  }
  __syncwarp();
  // ---- dWE2[j][k] += sum_p V[p][j] * E1[p][k] on the tensor cores (3xTF32) ----
  {
    const int g = lane >> 2, t = lane & 3, tx = t << 3;
#pragma unroll 1
    for (int ks = 0; ks < 4; ks++) {
      const float* Va = Vs + (ks * 8 + t) * ROWE;
      const float* Vb = Va + 4 * ROWE;
      const float* Ea = E1s + (ks * 8 + t) * ROWE;
      const float* Eb = Ea + 4 * ROWE;
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        split_tf32(Va[(mt * 16 + g) ^ tx], ah[mt][0], al[mt][0]);
        split_tf32(Va[(mt * 16 + g + 8) ^ tx], ah[mt][1], al[mt][1]);
        split_tf32(Vb[(mt * 16 + g) ^ tx], ah[mt][2], al[mt][2]);
        split_tf32(Vb[(mt * 16 + g + 8) ^ tx], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(Ea[(nt * 8 + g) ^ tx], bh0, bl0);
        split_tf32(Eb[(nt * 8 + g) ^ tx], bh1, bl1);
        mma_3xtf32(acc.cWE2[0][nt], ah[0], al[0], bh0, bh1, bl0, bl1);
        mma_3xtf32(acc.cWE2[1][nt], ah[1], al[1], bh0, bh1, bl0, bl1);
      }
    }
  }
  acc.s1 += colsum<ROWE>(Vs, lane);  // dbE2
  __syncwarp();                      // Vs rows may now be overwritten
  // ---- e1bar = WE2^T vbar, layer-1 reverse sweep; ubar_k -> Vs row ----
#pragma unroll 1
  for (int k4 = 0; k4 < NE; k4 += 4) {
    const float4 ev = LD4(&E1row[k4 ^ sx]);
    const float ea[4] = {ev.x, ev.y, ev.z, ev.w};
    float ub[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float* wrow = &w.WE2T[(k4 + i) * NE];
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int j4 = 0; j4 < NE; j4 += 8) {
        const float4 wa = LD4(&wrow[j4]), wb = LD4(&wrow[j4 + 4]);
        s0 = fmaf(wa.x, vbar[j4 + 0], s0); s0 = fmaf(wa.y, vbar[j4 + 1], s0);
        s0 = fmaf(wa.z, vbar[j4 + 2], s0); s0 = fmaf(wa.w, vbar[j4 + 3], s0);
        s1 = fmaf(wb.x, vbar[j4 + 4], s1); s1 = fmaf(wb.y, vbar[j4 + 5], s1);
        s1 = fmaf(wb.z, vbar[j4 + 6], s1); s1 = fmaf(wb.w, vbar[j4 + 7], s1);
      }
      ub[i] = (s0 + s1) * fmaf(-ea[i], ea[i], ea[i]);
    }
    ST4(&Vrow[k4 ^ sx], ub[0], ub[1], ub[2], ub[3]);
  }
  // ---- gate reverse sweep + dbE -> E1s row (free now) ----
  {
    float ch[32];
#pragma unroll
    for (int i = 0; i < NL; i++) {
      const float s = sigm(fmaf(R, w.WgL[i], w.bgL[i]));
      const float ub = gate_grads ? gbar * w.wg[i] * fmaf(-s, s, s) : 0.0f;
      ch[i] = ub * R;
      ch[10 + i] = ub;
      ch[20 + i] = gate_grads ? gbar * s : 0.0f;
    }
    ch[30] = gate_grads ? gbar : 0.0f;
    ch[31] = Ebar;
#pragma unroll
    for (int c4 = 0; c4 < 32; c4 += 4) ST4(&E1row[c4 ^ sx], ch[c4], ch[c4 + 1], ch[c4 + 2], ch[c4 + 3]);
  }
  __syncwarp();
  {
    float plain, weighted;
    colsum_w<ROWE>(Vs, lane, R, plain, weighted);
    acc.s2 += weighted;  // dWE1[k] = sum_p ubar_k R_p
    acc.s3 += plain;     // dbE1
  }
  acc.s4 += colsum<ROWE>(E1s, lane);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the fused kernel.  NEV = MLP evaluations per point (2: poc, 1: train.py); TRAIN = with reverse sweep
// ---------------------------------------------------------------------------------------------
template <int NEV>
__host__ __device__ constexpr int group_stash_floats() { return NEV * EVAL_STASH + ENET_STASH; }

template <int NEV, int G>
__host__ __device__ constexpr size_t step_smem_bytes() {
  return sizeof(Wts) + 16 /*mbarrier*/ + sizeof(float2) * G * 2 * 3 * 32 + sizeof(float) * G * group_stash_floats<NEV>();
}

template <int NEV, int G, bool TRAIN>
__global__ void __launch_bounds__((NEV + 1) * 32 * G, 1) pinn_step_kernel(const StepParams p) {
  constexpr int WPG = NEV + 1;  // warps per group
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Wts& w = *reinterpret_cast<Wts*>(smem_raw);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Wts));
  float2* mbox = reinterpret_cast<float2*>(smem_raw + sizeof(Wts) + 16);
  float* stash = reinterpret_cast<float*>(smem_raw + sizeof(Wts) + 16 + sizeof(float2) * G * 2 * 3 * 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = warp / WPG, role = warp % WPG;  // role < NEV: MLP evaluation; role == NEV: E-net + gate
  const bool is_mlp = role < NEV;
  const int sx = swz(lane);

  // ---- stage the weight image once per CTA: TMA bulk copy global -> shared, completion on an mbarrier ----
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)),
                 "r"((uint32_t)sizeof(Wts))
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&w)),
        "l"(p.wts), "r"((uint32_t)sizeof(Wts)), "r"(smem_u32(mbar))
        : "memory");
  }
  {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred q;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, q;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(mbar))
          : "memory");
    }
  }

  float* gstash = stash + grp * group_stash_floats<NEV>();
  float* Hs = gstash + (is_mlp ? role * EVAL_STASH : NEV * EVAL_STASH);
  float* Gs = Hs + (is_mlp ? 32 * ROWH : 32 * ROWE);  // for the E-net warp: Hs = E1s, Gs = Vs
  float2* gbox = mbox + grp * (2 * 3 * 32);

  double wpde = 0.0, wbc1 = 0.0, wbc2 = 0.0;
  if (TRAIN) { wpde = p.weights[0]; wbc1 = p.weights[1]; wbc2 = p.weights[2]; }
  const float w_pde = (float)wpde, w_bc1 = (float)wbc1, w_bc2 = (float)wbc2;
  const float sN = p.vc.sN, cL = p.vc.cL, cV = p.vc.cV, cE = p.vc.cE;

  MlpAcc macc;
  EnetAcc eacc;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) macc.cW2[i][j] = 0.0f;
  macc.s0 = macc.s1 = macc.s2 = 0.0f;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
      for (int c = 0; c < 4; c++) eacc.cWE2[a][b][c] = 0.0f;
  eacc.s0 = eacc.s1 = eacc.s2 = eacc.s3 = eacc.s4 = 0.0f;

  const long long ntiles = (p.n + 31) >> 5;
  const long long tstride = (long long)gridDim.x * G;
  int it = 0;
  for (long long tile = (long long)blockIdx.x * G + grp; tile < ntiles; tile += tstride, ++it) {
    const long long pidx = tile * 32 + lane;
    const bool valid = pidx < p.n;
    const long long pi = valid ? pidx : (p.n - 1);
    const Geom g = load_geom(p, pi);
    float2* box = gbox + (it & 1) * (3 * 32);

    // the evaluation at the inversion image swaps the roles of the two nuclei (poc/main.py:255-256)
    const bool sw = (role == 1) && is_mlp;
    const float a = sw ? g.f2 : g.f1, b = sw ? g.f1 : g.f2;
    const float al1 = sw ? g.al2 : g.al1, al2 = sw ? g.al1 : g.al2;
    const float al11 = sw ? g.al22 : g.al11, al22 = sw ? g.al11 : g.al22;
    const float al12 = g.al12;

    if (is_mlp) {
      float Nv, Dv;
      mlp_forward<TRAIN>(w, a, b, al1, al2, al11, al12, al22, Hs + lane * ROWH, Gs + lane * ROWH, sx, Nv, Dv);
      box[role * 32 + lane] = make_float2(Nv, Dv);
    } else {
      const float E = enet_forward<TRAIN>(w, g.R, Hs + lane * ROWE, Gs + lane * ROWE, sx);
      const float gt = gate_forward(w, g.R);
      box[2 * 32 + lane] = make_float2(E, gt);
    }
    named_barrier(1 + grp, WPG * 32);

    // ---- combine (every role recomputes the few scalars it needs) ----
    float2 m0 = box[lane];
    if (NEV == 2) { const float2 m1 = box[32 + lane]; m0.x += m1.x; m0.y += m1.y; }
    const float2 me = box[2 * 32 + lane];
    const float E = me.x, gate = me.y;
    const float N = fmaf(sN, m0.x, w.bo), DN = sN * m0.y;
    const float q = g.ir1 + g.ir2;
    const float fs = g.f1 + g.f2;
    const float psi = fmaf(gate, N, fs);
    const float lcao = fmaf(cL, fs, fmaf(cV - 2.0f * cL, fmaf(g.f1, g.ir1, g.f2 * g.ir2),
                                          cV * fmaf(g.f1, g.ir2, g.f2 * g.ir1)));
    const float inner = fmaf(cL, DN, cV * q * N);
    const float res = fmaf(gate, inner, fmaf(cE * E, psi, lcao));

    if (!TRAIN) {
      if (valid) {
        if (role == 0) {
          if (p.psi) p.psi[pidx] = psi;
          if (p.lap) p.lap[pidx] = g.al1 + g.al2 + gate * DN;
          if (p.res) p.res[pidx] = res;
          if (p.hpsi)
            p.hpsi[pidx] = fmaf(gate, fmaf(-0.5f, DN, -q * N),
                                fmaf(-0.5f, fs, -fmaf(g.f1, g.ir2, g.f2 * g.ir1)));
        } else if (!is_mlp) {
          if (p.E_out) p.E_out[pidx] = E;
        }
      }
      continue;
    }

    // ---- seeds of the reverse sweep (oracle/closed_form.py:loss_and_grad) ----
    float m1f, m2f;
    if (p.mask) {
      const unsigned mk = p.mask[pi];
      m1f = (mk & 1u) ? 1.0f : 0.0f;
      m2f = (mk & 2u) ? 1.0f : 0.0f;
    } else {
      m1f = (g.ir1 * p.bcut <= 1.0f) ? 1.0f : 0.0f;
      m2f = (g.ir2 * p.bcut <= 1.0f) ? 1.0f : 0.0f;
    }
    const float vw = valid ? 1.0f : 0.0f;
    const float rbar = 2.0f * w_pde * res * vw;
    const float pbar = 2.0f * fmaf(w_bc1, m1f, w_bc2 * m2f) * psi * vw;
    if (is_mlp) {
      const float lamN = fmaf(rbar, gate * fmaf(cV, q, cE * E), pbar * gate);
      const float lamD = rbar * cL * gate;
      float extra[8];
#pragma unroll
      for (int i = 0; i < 8; i++) extra[i] = 0.0f;
      if (role == 0) {
        extra[0] = res * res * vw;
        extra[1] = psi * psi * m1f * vw;
        extra[2] = psi * psi * m2f * vw;
        extra[3] = E * vw;
        extra[4] = p.base_grads ? lamN : 0.0f;  // dL/dbo
        extra[5] = m1f * vw;
        extra[6] = m2f * vw;
      }
      if (p.base_grads) {
        mlp_backward(w, a, b, al1, al2, al11, al12, al22, sN * lamN, sN * lamD, Hs, Gs, lane, extra, macc);
      } else if (role == 0) {
        mlp_extras_only(Hs, lane, extra, macc);
      }
    } else {
      if (valid && p.E_out) p.E_out[pidx] = E;
      const float gbar = fmaf(rbar, fmaf(cE * E, N, inner), pbar * N);
      const float Ebar = rbar * cE * psi;
      enet_backward(w, g.R, Ebar, gbar, p.gate_grads != 0, Hs, Gs, lane, eacc);
    }
  }

  if (!TRAIN) return;

  // ---- fold the per-warp accumulators into one row per CTA (fixed order -> deterministic) ----
  __syncthreads();
  double* red = reinterpret_cast<double*>(stash);  // NPART doubles, reuses the stash
  for (int i = tid; i < NPART; i += blockDim.x) red[i] = 0.0;
  __syncthreads();
  const int nwarps = WPG * G;
  for (int wv = 0; wv < nwarps; wv++) {
    if (warp == wv) {
      const int gq = lane >> 2, tq = lane & 3;
      if (is_mlp) {
#pragma unroll
        for (int nt = 0; nt < 2; nt++) {
          red[O_W2 + gq * NH + nt * 8 + 2 * tq] += macc.cW2[nt][0];
          red[O_W2 + gq * NH + nt * 8 + 2 * tq + 1] += macc.cW2[nt][1];
          red[O_W2 + (gq + 8) * NH + nt * 8 + 2 * tq] += macc.cW2[nt][2];
          red[O_W2 + (gq + 8) * NH + nt * 8 + 2 * tq + 1] += macc.cW2[nt][3];
        }
        red[lane < 16 ? O_B2 + lane : O_WO + lane - 16] += macc.s0;
        red[lane < 16 ? O_W1 + 2 * lane : O_W1 + 2 * (lane - 16) + 1] += macc.s1;
        if (lane < 16) red[O_B1 + lane] += macc.s2;
        else if (role == 0) {
          const int e = lane - 16;
          const int dst = e == 0 ? S_RES2 : e == 1 ? S_PSI1 : e == 2 ? S_PSI2 : e == 3 ? S_E
                        : e == 4 ? O_BO : e == 5 ? S_CNT1 : e == 6 ? S_CNT2 : -1;
          if (dst >= 0) red[dst] += macc.s2;
        }
      } else {
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int j = mt * 16 + gq, k = nt * 8 + 2 * tq;
            red[O_WE2 + j * NE + k] += eacc.cWE2[mt][nt][0];
            red[O_WE2 + j * NE + k + 1] += eacc.cWE2[mt][nt][1];
            red[O_WE2 + (j + 8) * NE + k] += eacc.cWE2[mt][nt][2];
            red[O_WE2 + (j + 8) * NE + k + 1] += eacc.cWE2[mt][nt][3];
          }
        red[O_WE + lane] += eacc.s0;
        red[O_BE2 + lane] += eacc.s1;
        red[O_WE1 + lane] += eacc.s2;
        red[O_BE1 + lane] += eacc.s3;
        const int dst = lane < 10 ? O_WGL + lane : lane < 20 ? O_BGL + lane - 10 : lane < 30 ? O_WG + lane - 20
                      : lane == 30 ? O_BG : O_BE;
        red[dst] += eacc.s4;
      }
    }
    __syncthreads();
  }
  double* row = p.partials + (size_t)blockIdx.x * NPART;
  for (int i = tid; i < NPART; i += blockDim.x) row[i] = red[i];
}

// ---------------------------------------------------------------------------------------------
// theta -> Wts image
// ---------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ th, Wts* __restrict__ out) {
  build_weight_image<true>(th, out, threadIdx.x, blockDim.x);
}

}  // namespace pinn

namespace pinn {

template <int NEV, int G, bool TRAIN>
static cudaError_t launch_step_t(const StepParams& p, int grid, cudaStream_t st) {
  auto kern = pinn_step_kernel<NEV, G, TRAIN>;
  constexpr size_t smem = step_smem_bytes<NEV, G>();
  static bool configured[16] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  kern<<<grid, (NEV + 1) * 32 * G, smem, st>>>(p);
  return cudaGetLastError();
}

constexpr int GROUPS = 4;

cudaError_t launch_step(int nev, bool train, const StepParams& p, int grid, cudaStream_t st) {
  if (nev == 2) return train ? launch_step_t<2, GROUPS, true>(p, grid, st) : launch_step_t<2, GROUPS, false>(p, grid, st);
  return train ? launch_step_t<1, GROUPS, true>(p, grid, st) : launch_step_t<1, GROUPS, false>(p, grid, st);
}

cudaError_t launch_prep(const float* theta, Wts* out, cudaStream_t st) {
  prep_weights_kernel<<<1, 256, 0, st>>>(theta, out);
  return cudaGetLastError();
}

}  // namespace pinn
#endif  // PINN_AB_BUILD
