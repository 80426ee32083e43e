// landscape.cuh -- landscape preparation of the experiment driver on the device
// (SURVEY.md 8(f) N2).
//
// What make_nsf does on the host for every landscape / trial
// (scripts/run_experiment.py:160-199 and navsim/util.pyx:76-88):
//   threshold     for_labeling = (V >= 200)                                          (:169)
//   modal filter  skimage.filters.rank.modal with a w x w footprint of ones on that 0/1
//                 image: the more frequent value inside the footprint clipped at the image
//                 border, 0 on a tie                                                 (:170-178)
//   grain labels  skimage.measure.label: 8-connected components of the non-zero pixels,
//                 numbered in raster order of their first pixel                      (:179)
//   grain areas   regionprops(...).equivalent_diameter = sqrt(4 area / pi)           (:180, :134)
//   chemistry     set_HS_where_equal: H, S of every labelled pixel from per-grain tables
//                 (util.pyx:76-88; the tables are drawn by the host RNG, :131-135)
// Here: one kernel each, on the planar landscape the engine already holds; the labels stay on
// the device and only the per-grain areas (a few KB) and tables cross the host link.
//
// Connected components: union-find in global memory.  Every foreground pixel starts as its own
// root (its linear index); one pass links it to its W, NW, N, NE foreground neighbours with an
// atomicMin-based union (the smaller root wins), so the final root of a component is its
// smallest linear index = its first pixel in raster order; a second pass flattens.  Roots are
// then numbered 1, 2, ... in raster order (per-row root counts, a host prefix sum over the
// rows, per-row numbering), which is exactly the reference's numbering.
#pragma once
#include "common.cuh"

__global__ void k_ls_threshold(const uint8_t *v_plane, int rows, int cols, int pitch, int threshold, uint8_t *mask)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < cols) mask[(size_t)y * cols + x] = v_plane[(size_t)y * pitch + x] >= threshold ? 1 : 0;
}

__global__ void k_ls_modal(const uint8_t *in, int rows, int cols, int w, uint8_t *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    // footprint of w x w ones centred on the pixel (w odd), clipped at the border
    const int h = w / 2;
    const int y0 = max(0, y - h), y1 = min(rows - 1, y + h), x0 = max(0, x - h), x1 = min(cols - 1, x + h);
    int ones = 0;
    for (int yy = y0; yy <= y1; yy++)
        for (int xx = x0; xx <= x1; xx++) ones += in[(size_t)yy * cols + xx];
    const int total = (y1 - y0 + 1) * (x1 - x0 + 1);
    out[(size_t)y * cols + x] = (2 * ones > total) ? 1 : 0;
}

__device__ __forceinline__ int nvb_uf_find(const int *parent, int i)
{
    int p = parent[i];
    while (p != i) { i = p; p = parent[i]; }
    return i;
}

__device__ __forceinline__ void nvb_uf_union(int *parent, int a, int b)
{
    while (true) {
        a = nvb_uf_find(parent, a);
        b = nvb_uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // a > b: hang the larger root under the smaller
        const int old = atomicMin(parent + a, b);
        if (old == a) return;                           // a was still a root: linked
        a = old;                                        // somebody re-parented a meanwhile: retry from there
    }
}

__global__ void k_ls_init(const uint8_t *mask, long long n, int *parent)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = mask[i] ? (int)i : -1;
}

__global__ void k_ls_link(const uint8_t *mask, int rows, int cols, int *parent)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int i = y * cols + x;
    if (!mask[i]) return;
    if (x > 0 && mask[i - 1]) nvb_uf_union(parent, i, i - 1);
    if (y > 0) {
        const int up = i - cols;
        if (mask[up]) nvb_uf_union(parent, i, up);
        if (x > 0 && mask[up - 1]) nvb_uf_union(parent, i, up - 1);
        if (x + 1 < cols && mask[up + 1]) nvb_uf_union(parent, i, up + 1);
    }
}

__global__ void k_ls_flatten(long long n, int *parent)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && parent[i] >= 0) parent[i] = nvb_uf_find(parent, (int)i);
}

// roots per row (a root is a pixel that is its own parent)
__global__ void k_ls_row_roots(const int *parent, int rows, int cols, int *row_count)
{
    const int y = blockIdx.x;
    int c = 0;
    for (int x = threadIdx.x; x < cols; x += blockDim.x) c += parent[(size_t)y * cols + x] == y * cols + x;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    __shared__ int s[32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w];
        row_count[y] = t;
    }
}

// number the roots of each row in column order, starting after the roots of the rows above:
// newid[root pixel] = 1-based label; one warp per row, ballot-based ranking
__global__ void k_ls_number(const int *parent, int rows, int cols, const int *row_offset, int *newid)
{
    const int y = blockIdx.x, lane = threadIdx.x;
    int base = row_offset[y];
    for (int x0 = 0; x0 < cols; x0 += 32) {
        const int x = x0 + lane;
        const bool root = x < cols && parent[(size_t)y * cols + x] == y * cols + x;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, root);
        if (root) newid[(size_t)y * cols + x] = base + __popc(m & ((1u << lane) - 1u)) + 1;
        base += __popc(m);
    }
}

__global__ void k_ls_relabel(const int *parent, const int *newid, long long n, long long *labels, int *area)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = parent[i];
    const int id = r >= 0 ? newid[r] : 0;
    labels[i] = id;
    if (id > 0) atomicAdd(area + id - 1, 1);
}

// set_HS_where_equal (util.pyx:76-88) on the engine's planar landscape: plane 0 = H, 1 = S
__global__ void k_ls_paint(const long long *labels, int rows, int cols, int pitch, long long plane_stride,
                           const uint8_t *H, const uint8_t *S, uint8_t *land)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const long long id = labels[(size_t)y * cols + x];
    if (id > 0) {
        land[(size_t)y * pitch + x] = H[id - 1];
        land[plane_stride + (size_t)y * pitch + x] = S[id - 1];
    }
}

// interleaved host layout (any strides) already uploaded as a contiguous [rows][cols][3] block
// -> planar landscape with optional flips (run_experiment.py:196-199)
__global__ void k_ls_planar(const uint8_t *hsv, int rows, int cols, int pitch, long long plane_stride, int flip_v,
                            int flip_h, uint8_t *land)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int sy = flip_v ? rows - 1 - y : y, sx = flip_h ? cols - 1 - x : x;
    const uint8_t *p = hsv + ((size_t)sy * cols + sx) * 3;
    land[(size_t)y * pitch + x] = p[0];
    land[plane_stride + (size_t)y * pitch + x] = p[1];
    land[2 * plane_stride + (size_t)y * pitch + x] = p[2];
}

__global__ void k_ls_flip_planes(const uint8_t *src, int rows, int cols, int pitch, long long plane_stride, int flip_v,
                                 int flip_h, uint8_t *dst)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, pl = blockIdx.z;
    if (x >= cols) return;
    const int sy = flip_v ? rows - 1 - y : y, sx = flip_h ? cols - 1 - x : x;
    dst[pl * plane_stride + (size_t)y * pitch + x] = src[pl * plane_stride + (size_t)sy * pitch + sx];
}

// ---- offline landscape generation: diffuse (navsim/util.pyx:186-235) --------------------------
// One explicit time step of the 2-D heat equation with periodic boundaries,
//   new[i][j] = m[i][j] + multiplier * (m[i+1][j] + m[i-1][j] - 4 m[i][j] + m[i][j+1] + m[i][j-1]),
// in the reference's operation order, no fused multiply-add: bit-identical to the Cython loop.
// HBM/L2 bound (16 bytes per point and step; two 2000^2 fields stay in L2).
__global__ void k_diffuse_step(const double *__restrict__ m, double *__restrict__ out, int side, double multiplier)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= side) return;
    const int ip = (i + 1 == side) ? 0 : i + 1, im = (i == 0) ? side - 1 : i - 1;
    const int jp = (j + 1 == side) ? 0 : j + 1, jm = (j == 0) ? side - 1 : j - 1;
    const double c = m[(size_t)i * side + j];
    double acc = __dadd_rn(m[(size_t)ip * side + j], m[(size_t)im * side + j]);   // util.pyx:218-219
    acc = __dsub_rn(acc, __dmul_rn(4.0, c));                                       // :220
    acc = __dadd_rn(acc, m[(size_t)i * side + jp]);                                // :221
    acc = __dadd_rn(acc, m[(size_t)i * side + jm]);                                // :222
    out[(size_t)i * side + j] = __dadd_rn(c, __dmul_rn(multiplier, acc));          // :216
}
