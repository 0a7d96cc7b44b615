/*
 * pinn_b200.h - C ABI of the B200-native PINN residual-and-gradient hot path.
 *
 * The reference (slitvinov/PINN_for_quantum_wavefunction_surfaces) has no FFI: its
 * hot path is Python (poc/main.py:341-355 NN_ion.LossFunctions, poc/main.py:321 +
 * 118 parametricPsi + hamiltonian, and the inline block train.py:41-57).  This
 * header is the boundary a binding would use instead (ctypes stub: INTEGRATION.md).
 * Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *  - every entry returns 0 on success, a positive cudaError_t value for CUDA
 *    failures and a negative PINN_E* code for argument errors; the message is
 *    available from pinn_last_error().  Nothing throws across this boundary.
 *  - all device buffers are owned by the caller; the library only owns the
 *    workspace allocated in pinn_create().  No allocation happens per call, so
 *    every *device* entry point is CUDA-graph capturable.
 *  - every call takes the stream explicitly (as a void* holding a cudaStream_t) and
 *    selects the handle's device itself, so it may be called from any host thread
 *    (torch runs autograd.Function.backward on its own thread).
 *  - theta is the packed parameter vector, 1521 scalars in NN_ion.state_dict() order
 *    with nn.Linear (out,in) layout (poc/main.py:233-245):
 *      W1(16,2) b1(16) W2(16,16) b2(16) wo(16) bo | WE1(32) bE1(32) WE2(32,32) bE2(32)
 *      wE(32) bE | WgL(10) bgL(10) wg(10) bg
 *    train.py's tuple (train.py:108-109) is the same set in (in,out) layout with the
 *    gate before the E-net; the host wrapper permutes/transposes.
 */
#ifndef PINN_B200_H
#define PINN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINN_THETA_SIZE 1521
#define PINN_NUM_TENSORS 16
#define PINN_SUMS_SIZE 8

/* variant: which formulation of the model/residual */
#define PINN_VARIANT_POC 0     /* poc/main.py: two MLP evaluations summed; res = -1/2 lap psi + V psi - E psi */
#define PINN_VARIANT_TRAINPY 1 /* train.py: one evaluation doubled; res = lap psi + (e + 1/r1 + 1/r2) psi */

/* dtype of the coordinate arrays x,y,z,R */
#define PINN_F32 0
#define PINN_F64 1

#define PINN_EINVAL (-1)  /* bad argument */
#define PINN_ENOTSUP (-2) /* device is not sm_100 or kernel image missing */

typedef struct pinn_handle pinn_handle;

/* Library / layout introspection. */
int pinn_version(void);
int pinn_theta_size(void);
/* Offsets (in scalars) of the 16 tensors inside theta; out must hold 17 ints (last = 1521). */
void pinn_theta_offsets(int* out);

/* Create a handle bound to CUDA device `device`; allocates the small workspace
 * (prepared weights, per-CTA partial sums, pinned staging for the *_host entry). */
int pinn_create(int device, pinn_handle** out);
int pinn_destroy(pinn_handle* h);
/* Message of the last failure on this handle (or of the last pinn_create failure when h==NULL). */
const char* pinn_last_error(pinn_handle* h);
/* Number of kernels this handle has launched so far (bench.py reports it as gpu_launches). */
int64_t pinn_launch_count(pinn_handle* h);

/* Which implementation of the fused step kernel runs (same results to fp32 rounding, same interface):
 *  PINN_ENGINE_TCGEN05 (default) - skinny mat-vecs as M=128 TF32 tcgen05.mma with activations in tensor memory (3xTF32);
 *  PINN_ENGINE_FFMA              - the same mat-vecs as FFMA chains (kept for A/B measurement).
 * The environment variable PINN_B200_ENGINE=ffma|tcgen05 selects the initial value at pinn_create. */
#define PINN_ENGINE_FFMA 0
#define PINN_ENGINE_TCGEN05 1
int pinn_set_engine(pinn_handle* h, int engine);
int pinn_get_engine(pinn_handle* h);

/* Optional timing of the fused step kernel alone (bench.py's roofline figure): between begin and collect
 * every pinn_loss_fwd_bwd brackets its step-kernel launch with CUDA events on the caller's stream;
 * collect synchronises them and returns the summed duration and the number of launches. */
int pinn_profile_begin(pinn_handle* h);
int pinn_profile_collect(pinn_handle* h, double* total_ms, int* launches);

/*
 * Training evaluation: replaces NN_ion.LossFunctions + Ltot.backward()
 * (poc/main.py:341-355, 403) and train.py:41-57 + 71.
 *
 *   Ltot = w[0] * sum_p res_p^2 + w[1] * sum_{p in set1} psi_p^2 + w[2] * sum_{p in set2} psi_p^2
 *
 * with the reference's means obtained for w = {1/n, 1/|set1|, 1/|set2|}; under
 * data-parallel sharding the caller passes the GLOBAL n and counts so that shard
 * results simply add.
 *
 *  x,y,z,R   device, n scalars each, dtype in_dtype (coordinates of the collocation points)
 *  mask      device, n bytes, bit0 = point in boundary set 1 (r1 >= BCcutoff), bit1 = set 2;
 *            NULL: the kernel derives the sets from r1,r2 >= bcutoff (fp32 compare)
 *  theta     device, 1521 float
 *  weights   device, 3 double {w_pde, w_bc1, w_bc2}; NULL: 1/n and 1/count are used
 *            (a small extra kernel counts the sets first)
 *  grad_mask bit i set = gradient of tensor i (state_dict order) is wanted; tensors
 *            whose bit is clear get zeros.  0x0FC0 is the reference's fine-tune mode
 *            (freezeBase + freezeDecayUnit, poc/main.py:305-319): base-MLP and gate
 *            reverse sweeps are skipped.
 *  sums      device out, 8 double: {Ltot, Lpde, Lbc, sum E, sum res^2, sum psi^2 set1,
 *            sum psi^2 set2, E of the last point}
 *  dtheta    device out, 1521 double: dLtot/dtheta
 *  E_out     device out or NULL, n float: E(R_p) (poc returns it, main.py:355; train.py prints its mean)
 */
int pinn_loss_fwd_bwd(pinn_handle* h, int variant, int64_t n,
                      const void* x, const void* y, const void* z, const void* R, int in_dtype,
                      const uint8_t* mask, const float* theta, const double* weights,
                      uint32_t grad_mask, float bcutoff,
                      double* sums, double* dtheta, float* E_out, void* stream);

/*
 * Inference: replaces the pair parametricPsi + hamiltonian (poc/main.py:321, 118;
 * call sites 451-454) and the residual of train.py:54.  Any output may be NULL.
 *  psi, lap (laplacian of psi), hpsi (-1/2 lap psi + V psi, Hartree form), res
 *  (the variant's PDE residual), E: device, n float each.
 */
int pinn_fields(pinn_handle* h, int variant, int64_t n,
                const void* x, const void* y, const void* z, const void* R, int in_dtype,
                const float* theta,
                float* psi, float* lap, float* hpsi, float* res, float* E, void* stream);

/*
 * Same as pinn_loss_fwd_bwd but with HOST buffers (the call a CPU-resident caller such
 * as the unmodified reference training loop makes): copies the coordinates host->device
 * in chunks on two streams so copies overlap the kernel, runs the step, copies the 8 sums
 * and 1521 gradients back and synchronises.  theta_host is 1521 double (the reference keeps
 * parameters in float64); weights_host is 3 double or NULL.  Pinned host memory is faster but
 * not required.
 */
int pinn_loss_fwd_bwd_host(pinn_handle* h, int variant, int64_t n,
                           const void* x, const void* y, const void* z, const void* R, int in_dtype,
                           const uint8_t* mask, const double* theta_host, const double* weights_host,
                           uint32_t grad_mask, float bcutoff,
                           double* sums_host, double* dtheta_host, float* E_out_host);

#ifdef __cplusplus
}
#endif
#endif /* PINN_B200_H */
