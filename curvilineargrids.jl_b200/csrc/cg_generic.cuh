// Host side of the GENERIC MappedGrid path: the caller's mapping (CUDA C++ source of `cg_map`, see
// cg_generic_prelude.h) is compiled with NVRTC for sm_100a, loaded as a module and launched through the driver API.
// libnvrtc and libcuda are resolved with dlopen at first use, so libcurvigrid_cuda.so keeps loading (and failing
// loudly on use) on a machine without a GPU.
//
// Reference seam replaced: MappedGrid takes arbitrary closures (/root/reference/src/grids/unified_grids/mapped_grid.jl:421-540)
// and differentiates them with ForwardDiff; the separable term lists of cg_mapped.cuh cover the analytic mappings of the
// reference's tests at roofline speed, this path covers everything else (about 100 mapping evaluations on 18-component
// jets per cell: a fallback, not a roofline kernel).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <string>
#include <vector>

#include "cg_generic_prelude.h"

namespace cg {

struct NvrtcApi {
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  bool ok = false;
  std::string why;
};
inline const NvrtcApi& nvrtc_api() {
  static const NvrtcApi api = [] {
    NvrtcApi a;
    void* h = nullptr;
    const char* e = getenv("CG_NVRTC_LIB");  // listed in include/curvigrid_cuda.h
    const char* names[] = {e, "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* n : names)
      if (n && !h) h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      a.why = "libnvrtc not found (set CG_NVRTC_LIB)";
      return a;
    }
    bool all = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) {
        all = false;
        a.why = std::string("missing NVRTC symbol ") + n;
      }
      return p;
    };
    a.CreateProgram = (decltype(a.CreateProgram))sym("nvrtcCreateProgram");
    a.CompileProgram = (decltype(a.CompileProgram))sym("nvrtcCompileProgram");
    a.GetProgramLogSize = (decltype(a.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    a.GetProgramLog = (decltype(a.GetProgramLog))sym("nvrtcGetProgramLog");
    a.GetCUBINSize = (decltype(a.GetCUBINSize))sym("nvrtcGetCUBINSize");
    a.GetCUBIN = (decltype(a.GetCUBIN))sym("nvrtcGetCUBIN");
    a.DestroyProgram = (decltype(a.DestroyProgram))sym("nvrtcDestroyProgram");
    a.GetErrorString = (decltype(a.GetErrorString))sym("nvrtcGetErrorString");
    a.ok = all;
    return a;
  }();
  return api;
}

struct CuApi {
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  bool ok = false;
  std::string why;
};
inline const CuApi& cu_api() {
  static const CuApi api = [] {
    CuApi a;
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      a.why = "libcuda.so.1 not found (no NVIDIA driver on this machine)";
      return a;
    }
    bool all = true;
    auto sym = [&](const char* n) {
      void* p = dlsym(h, n);
      if (!p) {
        all = false;
        a.why = std::string("missing driver symbol ") + n;
      }
      return p;
    };
    a.ModuleLoadData = (decltype(a.ModuleLoadData))sym("cuModuleLoadData");
    a.ModuleUnload = (decltype(a.ModuleUnload))sym("cuModuleUnload");
    a.ModuleGetFunction = (decltype(a.ModuleGetFunction))sym("cuModuleGetFunction");
    a.LaunchKernel = (decltype(a.LaunchKernel))sym("cuLaunchKernel");
    a.GetErrorString = (decltype(a.GetErrorString))sym("cuGetErrorString");
    a.ok = all;
    return a;
  }();
  return api;
}

// prelude + user source + kernels -> sm_100a cubin.  Returns "" on success, else the error text (NVRTC log included).
inline std::string generic_compile(const char* user_src, std::vector<char>& cubin) {
  const NvrtcApi& N = nvrtc_api();
  if (!N.ok) return N.why;
  std::string src = std::string(GENERIC_PRELUDE) + "\n#line 1 \"cg_map.cu\"\n" + user_src + "\n#line 1 \"cg_generic_kernels.cu\"\n" +
                    GENERIC_KERNELS;
  nvrtcProgram prog = nullptr;
  nvrtcResult r = N.CreateProgram(&prog, src.c_str(), "cg_generic.cu", 0, nullptr, nullptr);
  if (r != NVRTC_SUCCESS) return std::string("nvrtcCreateProgram: ") + N.GetErrorString(r);
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--device-int128"};
  r = N.CompileProgram(prog, 3, opts);
  std::string log;
  size_t ls = 0;
  if (N.GetProgramLogSize(prog, &ls) == NVRTC_SUCCESS && ls > 1) {
    log.resize(ls);
    N.GetProgramLog(prog, &log[0]);
  }
  if (r != NVRTC_SUCCESS) {
    N.DestroyProgram(&prog);
    return std::string("NVRTC: ") + N.GetErrorString(r) + "\n" + log;
  }
  size_t cs = 0;
  r = N.GetCUBINSize(prog, &cs);
  if (r != NVRTC_SUCCESS || cs == 0) {
    N.DestroyProgram(&prog);
    return "NVRTC produced no cubin";
  }
  cubin.resize(cs);
  r = N.GetCUBIN(prog, cubin.data());
  N.DestroyProgram(&prog);
  return r == NVRTC_SUCCESS ? "" : "nvrtcGetCUBIN failed";
}

// host mirrors of the parameter structs of the generic kernels (cg_generic_prelude.h)
struct GD {
  int dim, nhalo, scheme, what;
  long long nw[3], w0[3];
};
struct GOut {
  void *cell_fwd, *cell_inv, *face_fwd[3], *face_inv[3], *face_cons[3], *cjac, *cactive[3];
};
struct GCoord {
  void *node[3], *centroid[3], *face[3][3];
};

struct GenericMap {
  CUmodule mod = nullptr;
  CUfunction coords[2] = {nullptr, nullptr}, cell[2] = {nullptr, nullptr}, face[2] = {nullptr, nullptr};  // [f64, f32]
  double* d_params = nullptr;
  int nparams = 0;
  bool active = false;
  void destroy() {
    if (mod && cu_api().ok) cu_api().ModuleUnload(mod);
    if (d_params) cudaFree(d_params);
    *this = GenericMap();
  }
};

inline std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (cu_api().ok && cu_api().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUDA driver error " + std::to_string((int)r);
}

// compile + load; the device of the calling thread must be set (the runtime's primary context is used)
inline std::string generic_load(GenericMap& G, const char* user_src) {
  std::vector<char> cubin;
  std::string err = generic_compile(user_src, cubin);
  if (!err.empty()) return err;
  const CuApi& C = cu_api();
  if (!C.ok) return C.why;
  cudaFree(nullptr);  // make sure the primary context exists and is current
  CUmodule mod = nullptr;
  CUresult r = C.ModuleLoadData(&mod, cubin.data());
  if (r != CUDA_SUCCESS) return "cuModuleLoadData: " + cu_err(r);
  GenericMap M;
  M.mod = mod;
  const char* names[6] = {"cgg_coords_f64", "cgg_coords_f32", "cgg_cell_f64", "cgg_cell_f32", "cgg_face_f64", "cgg_face_f32"};
  CUfunction* dst[6] = {&M.coords[0], &M.coords[1], &M.cell[0], &M.cell[1], &M.face[0], &M.face[1]};
  for (int i = 0; i < 6; ++i) {
    r = C.ModuleGetFunction(dst[i], mod, names[i]);
    if (r != CUDA_SUCCESS) {
      C.ModuleUnload(mod);
      return std::string("cuModuleGetFunction(") + names[i] + "): " + cu_err(r);
    }
  }
  M.d_params = G.d_params;  // parameters survive a change of the source
  M.nparams = G.nparams;
  G.d_params = nullptr;
  G.destroy();
  G = M;
  G.active = true;
  return "";
}

inline std::string generic_launch(CUfunction fn, int64_t work, int threads, void** args, cudaStream_t st) {
  int64_t blocks = (work + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  CUresult r = cu_api().LaunchKernel(fn, (unsigned)blocks, 1, 1, (unsigned)threads, 1, 1, 0, (CUstream)st, args, nullptr);
  return r == CUDA_SUCCESS ? "" : "cuLaunchKernel: " + cu_err(r);
}

}  // namespace cg
