// ABI bookkeeping: version, error string, hash-grid geometry.
#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace ucsa {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UCSA_ERR_CUDA;
  }
  return UCSA_OK;
}

int set_max_dyn_smem(const void* fn, size_t bytes, const char* what) {
  if (bytes > 227 * 1024) {
    set_error("%s: needs %zu bytes of shared memory per CTA (limit 227 KB)", what, bytes);
    return UCSA_ERR_UNSUPPORTED;
  }
  const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu bytes): %s", what, bytes, cudaGetErrorString(e));
    return UCSA_ERR_CUDA;
  }
  return UCSA_OK;
}

}  // namespace ucsa

extern "C" {

int ucsa_abi_version(void) { return UCSA_ABI_VERSION; }

const char* ucsa_last_error_string(void) { return ucsa::g_error; }

// tcnn HashGrid geometry for the configuration at network_tcnn_semantics.py:36-46.  log2(per_level_scale) is
// taken in float32 (as tcnn stores it); exp2 is evaluated in double and rounded once to float32, so that the
// constants do not depend on the libm in use and every kernel (and the CPU oracle) sees identical values.
int ucsa_grid_desc_init(float per_level_scale, uint32_t base_resolution, uint32_t log2_hashmap_size,
                        ucsa_grid_desc* out) {
  UCSA_REQUIRE(out != nullptr, "grid_desc_init: null output");
  UCSA_REQUIRE(per_level_scale > 1.0f && base_resolution >= 2 && log2_hashmap_size >= 8 &&
                   log2_hashmap_size <= 24,
               "grid_desc_init: unsupported geometry");
  const float log2_pls = std::log2(per_level_scale);
  const uint64_t cap = 1ull << log2_hashmap_size;
  uint64_t off = 0;
  for (int l = 0; l < UCSA_GRID_LEVELS; ++l) {
    const float scale = static_cast<float>(
        std::exp2(static_cast<double>(l) * static_cast<double>(log2_pls)) * static_cast<double>(base_resolution) - 1.0);
    const uint32_t res = static_cast<uint32_t>(std::ceil(scale)) + 1u;
    const uint64_t cube = static_cast<uint64_t>(res) * res * res;
    uint64_t n = (cube + 7ull) / 8ull * 8ull;
    if (n > cap) n = cap;
    out->scale[l] = scale;
    out->res[l] = res;
    out->entries[l] = static_cast<uint32_t>(n);
    out->offset[l] = static_cast<uint32_t>(off);
    out->hashed[l] = cube > n ? 1u : 0u;
    off += n;
  }
  out->total_entries = static_cast<uint32_t>(off);
  return UCSA_OK;
}

}  // extern "C"
