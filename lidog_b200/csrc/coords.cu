// Voxelisation and coordinate maps: float32 quantisation, GPU open-addressing hash of packed
// voxel keys, first-occurrence unique / inverse maps, stride-s downsampling.
// Reference contract: ME.utils.sparse_quantize call sites (utils/datasets/semantickitti_bev.py:232-238),
// ME.SparseTensor construction (utils/pipelines/trainer_lighting_2d.py:151) and the stride-2
// coordinate maps implied by MinkowskiConvolution(stride=2) (utils/models/minkunet_bev.py:62-83).
// HBM-bound integer kernels: one coalesced pass over the rows per phase, table probes hit L2.
#include "common.cuh"
#include "runtime.cuh"
#include "scan.cuh"

namespace lg {

__global__ void __launch_bounds__(256)
    k_quantize(const float* __restrict__ pts, const int32_t* __restrict__ batch_of_row, int64_t n, float sx, float sy,
               float sz, int4* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // separately rounded IEEE division then floor: matches numpy float32 `floor(p / s)`
  float x = floorf(__fdiv_rn(pts[3 * i + 0], sx));
  float y = floorf(__fdiv_rn(pts[3 * i + 1], sy));
  float z = floorf(__fdiv_rn(pts[3 * i + 2], sz));
  int b = batch_of_row ? batch_of_row[i] : 0;
  out[i] = make_int4(b, (int)x, (int)y, (int)z);
}

// float64 points: the reference's TRAINING path.  Its augmentations multiply the float32 cloud by float64 matrices
// (utils/common/augmentation.py:10-20: `coords @ R` with R from scipy's expm), so sparse_quantize receives float64
// and numpy divides in float64; floor(x / 0.05) differs between the two precisions on cell boundaries.
__global__ void __launch_bounds__(256)
    k_quantize_f64(const double* __restrict__ pts, const int32_t* __restrict__ batch_of_row, int64_t n, double sx,
                   double sy, double sz, int4* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = floor(__ddiv_rn(pts[3 * i + 0], sx));
  const double y = floor(__ddiv_rn(pts[3 * i + 1], sy));
  const double z = floor(__ddiv_rn(pts[3 * i + 2], sz));
  const int b = batch_of_row ? batch_of_row[i] : 0;
  out[i] = make_int4(b, (int)x, (int)y, (int)z);
}

__device__ __forceinline__ int4 stride_coord(int4 c, int stride) {
  if (stride > 1) {
    c.y = floor_div(c.y, stride) * stride;
    c.z = floor_div(c.z, stride) * stride;
    c.w = floor_div(c.w, stride) * stride;
  }
  return c;
}

// Phase 1: insert every row; the slot keeps the MINIMUM row id (= first occurrence).
// (n_dev, when given, holds the true row count on the device -- the chained levels of lg_coords_pyramid never learn
// their sizes on the host; n is then only the upper bound the grid was sized for)
__global__ void __launch_bounds__(256)
    k_insert(const int4* __restrict__ coords, int64_t n, const int64_t* __restrict__ n_dev, int stride,
             HashSlot* __restrict__ table, unsigned long long mask, int* __restrict__ slot_of_row,
             int* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int4 c = stride_coord(coords[i], stride);
  if (!coord_in_range(c.x, c.y, c.z, c.w)) {
    *status = LG_ERR_RANGE;
    slot_of_row[i] = -1;
    return;
  }
  const unsigned long long key = pack_key(c.x, c.y, c.z, c.w);
  unsigned long long s = hash_key(key) & mask;
#pragma unroll 1
  for (;;) {
    unsigned long long prev = atomicCAS(&table[s].key, kEmptyKey, key);
    if (prev == kEmptyKey || prev == key) break;
    s = (s + 1) & mask;
  }
  atomicMin(&table[s].val, (unsigned)i);
  slot_of_row[i] = (int)s;
}

// Phase 2 (optional): a voxel whose points disagree on the label is marked.
__global__ void __launch_bounds__(256)
    k_label_disagree(const int* __restrict__ slot_of_row, const HashSlot* __restrict__ table,
                     const int32_t* __restrict__ labels, int64_t n, const int64_t* __restrict__ n_dev,
                     unsigned char* __restrict__ disagree) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int s = slot_of_row[i];
  if (s < 0) return;
  unsigned first = table[s].val;
  if (first != (unsigned)i && labels[i] != labels[first]) disagree[first] = 1;
}

struct FirstFlag {
  const int* slot_of_row;
  const HashSlot* table;
  const int64_t* n_dev;
  __device__ bool operator()(int64_t i) const {
    if (n_dev && i >= *n_dev) return false;
    int s = slot_of_row[i];
    return s >= 0 && table[s].val == (unsigned)i;
  }
};

struct FirstSink {
  const int4* coords;
  int stride;
  int4* out_coords;
  int64_t* unique_map;
  int* uid_of_first;  // [n], valid at first-occurrence rows
  const int32_t* labels;
  const unsigned char* disagree;
  int32_t ignore_label;
  int32_t* colabels;
  __device__ void operator()(int64_t i, int prefix, bool f) const {
    if (!f) return;
    out_coords[prefix] = stride_coord(coords[i], stride);
    unique_map[prefix] = i;
    uid_of_first[i] = prefix;
    if (colabels) colabels[prefix] = disagree[i] ? ignore_label : labels[i];
  }
};

// Phase 4: inverse map (reads the table while it still holds first-row ids).
__global__ void __launch_bounds__(256)
    k_inverse(const int* __restrict__ slot_of_row, const HashSlot* __restrict__ table,
              const int* __restrict__ uid_of_first, int64_t n, const int64_t* __restrict__ n_dev,
              int64_t* __restrict__ inverse_map) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = min(n, *n_dev);
  if (i >= n) return;
  int s = slot_of_row[i];
  inverse_map[i] = (s >= 0) ? (int64_t)uid_of_first[table[s].val] : -1;
}

// Phase 5: relabel the table from first-row id to unique id so later probes return rows of out_coords.
__global__ void __launch_bounds__(256)
    k_relabel(HashSlot* __restrict__ table, int64_t capacity, const int* __restrict__ uid_of_first) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= capacity) return;
  if (table[s].key != kEmptyKey) table[s].val = (unsigned)uid_of_first[table[s].val];
}

__global__ void k_write_status(int64_t* count_status, const int* status) { count_status[1] = (int64_t)*status; }

struct UniqueWorkspace {
  int* slot_of_row;
  int* uid_of_first;
  unsigned char* disagree;
  int* status;
  void* scan;
  size_t total;
};

static UniqueWorkspace carve_unique(void* base, int64_t n) {
  UniqueWorkspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? (char*)base + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  w.slot_of_row = (int*)take(sizeof(int) * (size_t)n);
  w.uid_of_first = (int*)take(sizeof(int) * (size_t)n);
  w.disagree = (unsigned char*)take((size_t)n);
  w.status = (int*)take(sizeof(int));
  w.scan = take(scan_workspace_bytes(n));
  w.total = off;
  return w;
}

}  // namespace lg

using namespace lg;

extern "C" int64_t lg_hash_capacity(int64_t n_keys) {
  int64_t c = 1024;
  while (c < 2 * n_keys) c <<= 1;
  return c;
}

extern "C" size_t lg_hash_bytes(int64_t capacity) { return (size_t)capacity * sizeof(HashSlot); }

extern "C" int lg_quantize_points(const float* points_xyz, const int32_t* batch_of_row, int64_t n, float size_x,
                                  float size_y, float size_z, int32_t* coords4_out, void* stream) {
  LG_CHECK_ARG(n >= 0 && size_x > 0 && size_y > 0 && size_z > 0, "lg_quantize_points: bad n or voxel size");
  if (n == 0) return LG_OK;
  LG_CHECK_ARG(points_xyz && coords4_out, "lg_quantize_points: null pointer");
  k_quantize<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(points_xyz, batch_of_row, n, size_x, size_y,
                                                                           size_z, (int4*)coords4_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" int lg_quantize_points_f64(const double* points_xyz, const int32_t* batch_of_row, int64_t n, double size_x,
                                      double size_y, double size_z, int32_t* coords4_out, void* stream) {
  LG_CHECK_ARG(n >= 0 && size_x > 0 && size_y > 0 && size_z > 0, "lg_quantize_points_f64: bad n or voxel size");
  if (n == 0) return LG_OK;
  LG_CHECK_ARG(points_xyz && coords4_out, "lg_quantize_points_f64: null pointer");
  k_quantize_f64<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(points_xyz, batch_of_row, n, size_x,
                                                                               size_y, size_z, (int4*)coords4_out);
  LG_LAUNCH_OK();
  return LG_OK;
}

extern "C" size_t lg_coords_unique_workspace(int64_t n) { return carve_unique(nullptr, n > 0 ? n : 1).total; }

namespace lg {
// One level: unique (strided) coordinates of `n` rows (n_dev: device-side count, n = upper bound) + its hash table.
static int unique_level(const int32_t* coords4, int64_t n, const int64_t* n_dev, int32_t stride, void* table,
                        int64_t capacity, int32_t* out_coords4, int64_t* unique_map, int64_t* inverse_map,
                        const int32_t* labels, int32_t ignore_label, int32_t* colabels, int64_t* count_status,
                        void* workspace, cudaStream_t stream) {
  LG_CUDA_OK(cudaMemsetAsync(table, 0xFF, lg_hash_bytes(capacity), stream));
  LG_CUDA_OK(cudaMemsetAsync(count_status, 0, 2 * sizeof(int64_t), stream));
  if (n == 0) return LG_OK;
  UniqueWorkspace w = carve_unique(workspace, n);
  LG_CUDA_OK(cudaMemsetAsync(w.status, 0, sizeof(int), stream));
  const unsigned grid = (unsigned)ceil_div(n, 256);
  HashSlot* tab = (HashSlot*)table;
  k_insert<<<grid, 256, 0, stream>>>((const int4*)coords4, n, n_dev, stride, tab, (unsigned long long)(capacity - 1),
                                     w.slot_of_row, w.status);
  LG_LAUNCH_OK();
  if (colabels) {
    LG_CUDA_OK(cudaMemsetAsync(w.disagree, 0, (size_t)n, stream));
    k_label_disagree<<<grid, 256, 0, stream>>>(w.slot_of_row, tab, labels, n, n_dev, w.disagree);
    LG_LAUNCH_OK();
  }
  FirstFlag flag{w.slot_of_row, tab, n_dev};
  FirstSink sink{(const int4*)coords4, stride,  (int4*)out_coords4, unique_map, w.uid_of_first,
                 labels,               w.disagree, ignore_label,       colabels};
  int rc = device_scan(flag, sink, n, count_status, w.scan, stream);
  if (rc != LG_OK) return rc;
  k_inverse<<<grid, 256, 0, stream>>>(w.slot_of_row, tab, w.uid_of_first, n, n_dev, inverse_map);
  LG_LAUNCH_OK();
  k_relabel<<<(unsigned)ceil_div(capacity, 256), 256, 0, stream>>>(tab, capacity, w.uid_of_first);
  LG_LAUNCH_OK();
  k_write_status<<<1, 1, 0, stream>>>(count_status, w.status);
  LG_LAUNCH_OK();
  return LG_OK;
}
}  // namespace lg

extern "C" int lg_coords_unique(const int32_t* coords4, int64_t n, int32_t stride, void* table, int64_t capacity,
                                int32_t* out_coords4, int64_t* unique_map, int64_t* inverse_map,
                                const int32_t* labels, int32_t ignore_label, int32_t* colabels, int64_t* count_status,
                                void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LG_CHECK_ARG(n >= 0 && stride >= 1, "lg_coords_unique: bad n/stride");
  LG_CHECK_ARG(capacity >= 2 * n && (capacity & (capacity - 1)) == 0 && capacity >= 1024,
               "lg_coords_unique: capacity must be a power of two >= max(1024, 2n)");
  LG_CHECK_ARG(table && count_status, "lg_coords_unique: null table/count_status");
  LG_CHECK_ARG((colabels == nullptr) || labels, "lg_coords_unique: colabels requested without labels");
  if (n > 0) {
    LG_CHECK_ARG(coords4 && out_coords4 && unique_map && inverse_map && workspace, "lg_coords_unique: null pointer");
    const size_t need = carve_unique(nullptr, n).total;
    LG_CHECK_ARG(workspace_bytes >= need, "lg_coords_unique: workspace too small (%zu < %zu)", workspace_bytes, need);
  }
  return unique_level(coords4, n, nullptr, stride, table, capacity, out_coords4, unique_map, inverse_map, labels,
                      ignore_label, colabels, count_status, workspace, stream);
}

/* Every coordinate level of a batch in ONE call and without a host round trip between the levels (see the header). */
extern "C" int lg_coords_pyramid(const int32_t* coords4, int64_t n, const int32_t* labels, int32_t ignore_label,
                                 int32_t* colabels, int32_t n_levels, const int32_t* strides, const lgLevelOut* levels,
                                 int64_t* counts_dev, int64_t* counts_host, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LG_CHECK_ARG(n >= 1 && n_levels >= 1 && n_levels <= 8 && strides && levels && counts_dev && coords4,
               "lg_coords_pyramid: bad arguments");
  LG_CHECK_ARG((colabels == nullptr) || labels, "lg_coords_pyramid: colabels requested without labels");
  for (int l = 0; l < n_levels; ++l) {
    const lgLevelOut& L = levels[l];
    LG_CHECK_ARG(strides[l] >= 1 && (l == 0 || strides[l] > strides[l - 1]), "lg_coords_pyramid: strides must increase");
    LG_CHECK_ARG(L.table && L.coords4 && L.unique_map && L.inverse_map && L.capacity >= 2 * n &&
                     (L.capacity & (L.capacity - 1)) == 0 && L.capacity >= 1024,
                 "lg_coords_pyramid: level %d needs table / outputs for the upper bound of %lld rows", l, (long long)n);
  }
  ArenaCursor ar;
  const size_t ws = carve_unique(nullptr, n).total;
  int rc = arena_begin(stream, ws, &ar);
  if (rc) return rc;
  void* workspace = arena_take(&ar, ws);
  for (int l = 0; l < n_levels; ++l) {
    const lgLevelOut& L = levels[l];
    rc = unique_level(l == 0 ? coords4 : levels[l - 1].coords4, n, l == 0 ? nullptr : counts_dev + 2 * (l - 1),
                      strides[l], L.table, L.capacity, L.coords4, L.unique_map, L.inverse_map, l == 0 ? labels : nullptr,
                      ignore_label, l == 0 ? colabels : nullptr, counts_dev + 2 * l, workspace, stream);
    if (rc) return rc;
  }
  if (counts_host)
    LG_CUDA_OK(cudaMemcpyAsync(counts_host, counts_dev, sizeof(int64_t) * 2 * n_levels, cudaMemcpyDeviceToHost, stream));
  return LG_OK;
}
