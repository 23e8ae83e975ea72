// The connections of one z-marching column tile, staged in shared memory before the march.
//
// The tiled kernels visit the planes of a tile in ascending order and need, per plane, the connections that sit in the
// tile's cells (scatter of the rates into the divergence, welldata_processor.py:170-224; inner-boundary error,
// physics_loss.py:189).  Searching the cell-sorted table once per cell and plane costs eleven DEPENDENT global loads
// per search on a 2048-connection lattice; the threads that own a well column then hold their whole CTA at every plane
// barrier (measured on B200, BASELINE config 5: 29 % of the step).  Here the tile's columns are found once,
// their layer-sorted connection lists (SrmDev::col_*) and the per-sample value each connection contributes (rate or
// d rate / d p) are copied to shared memory, and the march advances one cursor per well cell: no global load and no
// search inside the plane loop.
#pragma once
#include "srm_internal.cuh"

namespace {

struct WellColsDev {         // the column lists of SrmDev, for kernels that take a slim parameter block
  int32_t n_cols;
  const int32_t* col_rem;
  const int32_t* col_ptr;
  const int2* col_ent;
};
__host__ __device__ __forceinline__ WellColsDev well_cols_of(const SrmDev& P) { return WellColsDev{P.n_cols, P.col_rem, P.col_ptr, P.col_ent}; }

// NT threads, a tile of TW x TY cells, at most MAXSLOT well cells and CAP connections per tile (more: `overflow`, the
// caller keeps its search path)
template <int NT, int TW, int TY, int MAXSLOT, int CAP>
struct WellTile {
  int32_t n_slots, n_ent;
  int32_t beg[MAXSLOT], end[MAXSLOT], pos[MAXSLOT], src[MAXSLOT];
  int32_t w[CAP];             // position in the cell-sorted table (index into the per-sample qw / dqdp / divqw rows)
  float val[CAP];             // the per-sample value of the connection
  uint16_t lay[CAP];
  unsigned char slot_of[TY * TW];   // 0: no connection in this column, else slot + 1

  // block-wide; returns true when the tile holds a connection.  `overflow` (block-uniform) says the lists did not fit.
  // Ends with a barrier: slot_of / lists are visible to every thread.
  __device__ __forceinline__ bool build(const WellColsDev& C, int W, int D, int x0, int y0, const float* __restrict__ vals /* row of this sample, or null */,
                                        bool& overflow) {
    const int tid = threadIdx.x;
    for (int i = tid; i < TY * TW; i += NT) slot_of[i] = 0;
    if (tid == 0) { n_slots = 0; n_ent = 0; }
    __syncthreads();
    for (int c = tid; c < C.n_cols; c += NT) {
      const int rem = C.col_rem[c];
      const int j = rem / W, i = rem - j * W;
      if (i >= x0 && i < x0 + TW && j >= y0 && j < y0 + TY) {
        const int s = atomicAdd(&n_slots, 1);
        if (s < MAXSLOT) {
          const int first = C.col_ptr[c], cnt = C.col_ptr[c + 1] - first;
          const int off = atomicAdd(&n_ent, cnt);
          beg[s] = off; end[s] = off + cnt; pos[s] = off; src[s] = first;
          slot_of[(j - y0) * TW + (i - x0)] = (unsigned char)(s + 1);
        }
      }
    }
    __syncthreads();
    const int ns = n_slots, ne = n_ent;
    overflow = ns > MAXSLOT || ne > CAP || D > 65535;
    if (ns > 0 && !overflow) {
      for (int s = 0; s < ns; ++s) {
        const int b0 = beg[s], cnt = end[s] - b0, from = src[s];
        for (int t = tid; t < cnt; t += NT) {
          const int2 e = C.col_ent[from + t];
          lay[b0 + t] = (uint16_t)e.x;
          w[b0 + t] = e.y;
          if (vals) val[b0 + t] = vals[e.y];
        }
      }
    }
    __syncthreads();
    return ns > 0;
  }

  // slots of CPT x-adjacent cells starting at local cell `local`, one byte each (0: none)
  template <int CPT>
  __device__ __forceinline__ uint32_t slots_of(int local) const {
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) v |= (uint32_t)slot_of[local + c] << (8 * c);
    return v;
  }

  // the connections of slot `s` (slot byte - 1) in layer m: [first, last) in w / val; advances the slot's cursor.
  // Only the thread that owns the cell calls this, once per plane, planes ascending.
  __device__ __forceinline__ void take(int s, int m, int& first, int& last) {
    int p = pos[s];
    const int e = end[s];
    first = p;
    while (p < e && lay[p] == (uint16_t)m) ++p;
    last = p;
    pos[s] = p;
  }
};

}  // namespace
