// Drop-in for the reference's astaroth/kernels.cu: the four functions its driver (astaroth/astaroth.cu) calls --
//   integrate_substep            astaroth/kernels.cu:62-87   (launches solve<step>, block (32,1,4))
//   acDeviceLoadDefaultUniforms  astaroth/kernels.cu:190-216
//   acDeviceLoadMeshInfo         astaroth/kernels.cu:165-188
//   acDeviceLoadScalarUniform    astaroth/kernels.cu:89-106
// implemented over sb_astaroth_substep (stencil_b200/csrc/astaroth.cu).  Compiled only for the drop-in driver build
// (`make drivers` -> bin/astaroth_b200) with the reference's own astaroth/*.h on the include path, so the declarations,
// enums and VertexBufferArray are the reference's, unmodified; nothing of astaroth/ is copied into this repository.
//
// The reference keeps the uniforms in a __constant__ AcMeshInfo per device and updates it with stream-ordered
// cudaMemcpyToSymbolAsync; here they live in a host-side table per device and travel as kernel parameters, which is
// equivalent for the driver's usage (a uniform is set before the launches that read it, from the same host thread).
#include "kernels.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "stencil_b200.h"

namespace {

constexpr int kMaxDevices = 64;
AcMeshInfo g_info[kMaxDevices];
bool g_init[kMaxDevices];
std::mutex g_mu;

AcMeshInfo &info_of(int device) {
  if (device < 0 || device >= kMaxDevices) {
    fprintf(stderr, "astaroth_kernels: device %d out of range\n", device);
    exit(EXIT_FAILURE);
  }
  if (!g_init[device]) {
    memset(&g_info[device], 0, sizeof(AcMeshInfo));
    g_init[device] = true;
  }
  return g_info[device];
}

bool valid(double v) { return !std::isnan(v) && !std::isinf(v); }

// the *_DEFAULT_VALUE statics of astaroth/user_kernels.h:30-35, 329, 367 (all others are zero-initialised statics)
void load_defaults(AcMeshInfo &m) {
  memset(&m, 0, sizeof(m));
  const double ds = 0.04908738521;
  m.real_params[AC_dsx] = m.real_params[AC_dsy] = m.real_params[AC_dsz] = ds;
  m.real_params[AC_inv_dsx] = m.real_params[AC_inv_dsy] = m.real_params[AC_inv_dsz] = 1.0 / ds;
  m.real_params[AC_cs_sound] = 1.0;
  m.real_params[AC_cs2_sound] = 1.0;
}

} // namespace

extern "C" {

AcResult acDeviceLoadDefaultUniforms(const int device) {
  std::lock_guard<std::mutex> lock(g_mu);
  load_defaults(info_of(device));
  return AC_SUCCESS;
}

AcResult acDeviceLoadScalarUniform(const int device, cudaStream_t /*stream*/, const AcRealParam param, const AcReal value) {
  if (param < 0 || param >= NUM_REAL_PARAMS) {
    fprintf(stderr, "WARNING: invalid AcRealParam %d.\n", param);
    return AC_FAILURE;
  }
  if (!valid(value)) { // the NaN-filled entries of a parsed config leave the defaults in place (astaroth/kernels.cu:96-100)
    fprintf(stderr, "WARNING: Passed an invalid value %g to device constant %s. Skipping.\n", (double)value, realparam_names[param]);
    return AC_FAILURE;
  }
  std::lock_guard<std::mutex> lock(g_mu);
  info_of(device).real_params[param] = value;
  return AC_SUCCESS;
}

AcResult acDeviceLoadMeshInfo(const int device, const AcMeshInfo meshInfo) {
  {
    std::lock_guard<std::mutex> lock(g_mu);
    AcMeshInfo &m = info_of(device);
    for (int i = 0; i < NUM_INT_PARAMS; ++i) m.int_params[i] = meshInfo.int_params[i];
    for (int i = 0; i < NUM_INT3_PARAMS; ++i) m.int3_params[i] = meshInfo.int3_params[i];
  }
  for (int i = 0; i < NUM_REAL_PARAMS; ++i) acDeviceLoadScalarUniform(device, 0, (AcRealParam)i, meshInfo.real_params[i]);
  return AC_SUCCESS;
}

AcResult integrate_substep(const int stepNumber, cudaStream_t stream, Rect3 cr, VertexBufferArray vba) {
  int device = 0;
  CUDA_RUNTIME(cudaGetDevice(&device)); // the driver calls d.set_device() first (astaroth/astaroth.cu:568)
  sb_astaroth_params p;
  int64_t raw[3];
  {
    std::lock_guard<std::mutex> lock(g_mu);
    const AcMeshInfo &m = info_of(device);
    p.inv_dsx = m.real_params[AC_inv_dsx], p.inv_dsy = m.real_params[AC_inv_dsy], p.inv_dsz = m.real_params[AC_inv_dsz];
    p.dt = m.real_params[AC_dt];
    p.cs2_sound = m.real_params[AC_cs2_sound], p.gamma = m.real_params[AC_gamma], p.cp_sound = m.real_params[AC_cp_sound];
    p.lnrho0 = m.real_params[AC_lnrho0], p.lnT0 = m.real_params[AC_lnT0];
    p.mu0 = m.real_params[AC_mu0], p.nu_visc = m.real_params[AC_nu_visc], p.zeta = m.real_params[AC_zeta], p.eta = m.real_params[AC_eta];
    raw[0] = m.int_params[AC_mx], raw[1] = m.int_params[AC_my], raw[2] = m.int_params[AC_mz];
  }
  const void *in[NUM_VTXBUF_HANDLES];
  void *out[NUM_VTXBUF_HANDLES];
  for (int i = 0; i < NUM_VTXBUF_HANDLES; ++i) in[i] = vba.in[i], out[i] = vba.out[i];
  const int64_t lo[3] = {int64_t(cr.lo.x), int64_t(cr.lo.y), int64_t(cr.lo.z)};
  const int64_t hi[3] = {int64_t(cr.hi.x), int64_t(cr.hi.y), int64_t(cr.hi.z)};
  if (sb_astaroth_substep(stepNumber, in, out, int(sizeof(AcReal)), raw, lo, hi, &p, 0, stream) != SB_OK) {
    fprintf(stderr, "integrate_substep: %s\n", sb_last_error());
    exit(EXIT_FAILURE); // the reference's error convention: print + exit
  }
  return AC_SUCCESS;
}

} // extern "C"
