// Peer-store bandwidth probe: how much of the NVLink write bandwidth do warp stores reach, against the size of the
// contiguous segment one warp instruction writes?  (The fused delivery of the rasteriser writes 32-pixel tile rows:
// 128-byte segments, the next row 512 bytes further.)   Built by tools/p2p_store_probe.py.
#include <cuda_runtime.h>
#include <cstdint>
extern "C" {
// seg_floats contiguous floats per warp instruction (32 = one float per lane, 64 = float2, 128 = float4);
// consecutive instructions of a warp walk down `rows` rows that are `pitch_floats` apart (tile-row pattern).
__global__ void store_probe(float* __restrict__ dst, const size_t n_floats, const int vec, const int pitch_floats, const int rows) {
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int seg = 32 * vec;                       // floats per instruction
  const size_t block_floats = (size_t)pitch_floats * rows;   // a "tile column" of rows x pitch
  const int segs_per_row = pitch_floats / seg;
  const size_t n_units = n_floats / block_floats * segs_per_row;
  for (size_t u = warp; u < n_units; u += n_warps) {
    const size_t blk = u / segs_per_row, s = u % segs_per_row;
    float* base = dst + blk * block_floats + s * seg + lane * vec;
    for (int r = 0; r < rows; ++r) {
      float* q = base + (size_t)r * pitch_floats;
      if (vec == 1) *q = 1.0f;
      else if (vec == 2) *(float2*)q = make_float2(1.f, 1.f);
      else *(float4*)q = make_float4(1.f, 1.f, 1.f, 1.f);
    }
  }
}
int probe_launch(float* dst, size_t n_floats, int vec, int pitch_floats, int rows, int blocks, cudaStream_t st) {
  store_probe<<<blocks, 256, 0, st>>>(dst, n_floats, vec, pitch_floats, rows);
  return (int)cudaGetLastError();
}
}
