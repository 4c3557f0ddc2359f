// Packed layout of W = L^-1 shared by its writer (gp_fit.cu: pcabo_gp_pack_w) and its reader (lbfgsb_eval.cuh).
#pragma once

namespace lbeval {

// Written by pcabo_gp_pack_w, read by the optimiser kernel with ONE bulk copy per panel.
// The lower triangle is cut in groups of 8 rows; group g holds rows 8 g .. 8 g + 7, columns 0 .. 8 g + 7 (zero right of
// the diagonal and below row n), row-major with the padded row stride grp_rs(g), and the groups follow each other
// without gaps. A bulk copy costs the copy engine ~70-90 cycles whatever its size (tools/micro/bulk_bw.cu: 16 row
// copies of 3.4 KB reach 36 B/clk per SM, one copy of 55 KB reaches 100), hence a layout in which a panel of R rows
// (R / 8 consecutive groups) is one contiguous piece of memory.
__host__ __device__ inline int grp_rs(int g) { return 8 * g + ((g & 1) ? 12 : 20); }        // = pad4mod16(8 g + 8)
__host__ __device__ inline long long grp_off(int g) {                                        // doubles before group g
  return 32LL * g * (g - 1) + 160LL * ((g + 1) / 2) + 96LL * (g / 2);
}
__host__ __device__ inline int ngroups(int n) { return (n + 7) >> 3; }
__host__ __device__ inline long long packed_doubles(int n) { return grp_off(ngroups(n)); }

}  // namespace lbeval
