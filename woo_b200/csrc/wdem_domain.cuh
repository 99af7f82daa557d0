// wdem_domain.cuh — kernels of the 1-D slab decomposition: ownership, migration of spheres together with the history
// rows of their real contacts, ghost export at collider full runs, and the per-step ghost refresh (pos, vel, angVel).
// Cross-slab pairs are evaluated on both sides from bit-identical inputs (ghost = bit copy, history created in the same step on both
// sides), so both replicas of a contact stay bit-identical; forces are reduced only into owned spheres — no reverse exchange.
#pragma once
#include "wdem_kernels.cuh"

namespace wd {

#define WDEM_MAX_RANKS 16

struct DomainU {
  int rank, world, axis, periodic;
  double bounds[WDEM_MAX_RANKS + 1];   // slab r = [bounds[r], bounds[r+1]) in the slab coordinate (periodic: wrapped unsheared coordinate)
  double reach;                        // (largest half-extent + verletDist): how far an owned box of any rank sticks out of its slab
  double vd;
  double dim;                          // periodic: cell size along the slab axis
  PeriU pu;
};

__device__ __forceinline__ double slabCoord(const DomainU& D, double4 p) {
  if (!D.periodic) return comp(xyz(p), D.axis);
  const d3 u = mulMV(D.pu.unshear, xyz(p));
  return wrap01(comp(u, D.axis), D.dim);
}
__device__ __forceinline__ int ownerOf(const DomainU& D, double x) {
  int r = 0;
  while (r + 1 < D.world && x >= D.bounds[r + 1]) r++;
  return r;
}
// half-extent of the sphere's enlarged AABB along the slab axis
__device__ __forceinline__ double slabHalf(const DomainU& D, double4 p) {
  if (!D.periodic) return p.w + D.vd;
  d3 mn, mx; periAabb(D.pu, p, D.vd, mn, mx);
  return .5 * (comp(mx, D.axis) - comp(mn, D.axis));
}

// key 0: stays owned; 1+r: emigrates to rank r; world+1: ghost of the previous generation (dropped)
__global__ void __launch_bounds__(256) k_dom_classify(Spheres S, int n, DomainU D, uint32_t* keys, int* slots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t key;
  if (S.flags[i] & F_GHOST) key = D.world + 1;
  else { const int o = ownerOf(D, slabCoord(D, S.posr[i])); key = (o == D.rank) ? 0u : (uint32_t)(1 + o); }
  keys[i] = key; slots[i] = i;
}
// first index of every key in the sorted key array (start[] pre-filled with n)
__global__ void __launch_bounds__(256) k_dom_starts(const uint32_t* __restrict__ keys, int n, int* start) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || keys[i] != keys[i - 1]) start[keys[i]] = i;
}
__global__ void k_fill_int(int* p, int n, int v) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

struct PRec {   // a migrating sphere: Node + DemData + Sphere + Particle fields
  double4 posr, velm, angi, ori; double f[6]; uint32_t flags, mask; int32_t mat, id, up, pad;
};
struct GRec {   // a ghost: what the contact functors read
  double4 posr, velm, angi; uint32_t flags, mask; int32_t mat, id;
};
struct CRec {   // history row of a real contact keyed by particle ids
  int32_t idA, idB; uint32_t flags; int32_t stepCreated; int32_t cd[3]; int32_t pad; int32_t pairY, pad2;
  double t[6]; double4 cp; double fl[3]; double ph[5]; double tq[3]; double4 Fg;
};

__global__ void __launch_bounds__(256) k_dom_pack_particles(Spheres S, const int* __restrict__ slots, int first, int cnt, PRec* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int i = slots[first + k];
  PRec r;
  r.posr = S.posr[i]; r.velm = S.velm[i]; r.angi = S.angi[i]; r.ori = S.ori[i];
  r.f[0] = S.fx[i]; r.f[1] = S.fy[i]; r.f[2] = S.fz[i]; r.f[3] = S.tx[i]; r.f[4] = S.ty[i]; r.f[5] = S.tz[i];
  r.flags = S.flags[i]; r.mask = S.mask[i]; r.mat = S.mat[i]; r.id = S.id[i]; r.up = S.up[i]; r.pad = 0;
  out[k] = r;
}
__global__ void __launch_bounds__(256) k_dom_unpack_particles(Spheres S, int first, int cnt, const PRec* __restrict__ in) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int i = first + k;
  const PRec r = in[k];
  S.posr[i] = r.posr; S.velm[i] = r.velm; S.angi[i] = r.angi; S.ori[i] = r.ori;
  S.fx[i] = r.f[0]; S.fy[i] = r.f[1]; S.fz[i] = r.f[2]; S.tx[i] = r.f[3]; S.ty[i] = r.f[4]; S.tz[i] = r.f[5];
  S.lx[i] = r.posr.x; S.ly[i] = r.posr.y; S.lz[i] = r.posr.z;
  S.flags[i] = (r.flags & ~(F_GHOST | F_ARRIVED)) | F_ARRIVED; S.mask[i] = r.mask; S.mat[i] = r.mat; S.id[i] = r.id; S.up[i] = r.up;
}
// kept spheres: gather in (stable) slot order into the new generation
__global__ void __launch_bounds__(256) k_dom_gather_kept(Spheres dst, Spheres src, const int* __restrict__ slots, int cnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int o = slots[k];
  dst.posr[k] = src.posr[o]; dst.velm[k] = src.velm[o]; dst.angi[k] = src.angi[o]; dst.ori[k] = src.ori[o];
  dst.fx[k] = src.fx[o]; dst.fy[k] = src.fy[o]; dst.fz[k] = src.fz[o]; dst.tx[k] = src.tx[o]; dst.ty[k] = src.ty[o]; dst.tz[k] = src.tz[o];
  dst.lx[k] = src.lx[o]; dst.ly[k] = src.ly[o]; dst.lz[k] = src.lz[o];
  dst.flags[k] = src.flags[o] & ~(F_GHOST | F_ARRIVED); dst.mask[k] = src.mask[o]; dst.mat[k] = src.mat[o]; dst.id[k] = src.id[o]; dst.up[k] = src.up[o];
}

// history rows of emigrants: count, then pack (one thread per emigrant walks its incidence row of the previous table)
__global__ void __launch_bounds__(128) k_dom_count_rows(const int* __restrict__ slots, int first, int cnt, OldCsr oc, const uint32_t* __restrict__ oflags, int* rowCnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  int c = 0;
  if (oc.valid) { const int s = slots[first + k]; for (int e = oc.rowStart[s]; e < oc.rowStart[s + 1]; e++) { const uint32_t f = oflags[oc.rowEntry[e] >> 1]; if ((f & CF_REAL) && !(f & CF_DEAD)) c++; } }
  rowCnt[k] = c;
}
// the same for ALL spheres in destination order (sorted keys): 0 for everything that does not emigrate, so that an exclusive scan over the whole array gives the
// row offsets of the emigrants without the host having to know how many there are
__global__ void __launch_bounds__(128) k_dom_count_rows_all(const uint32_t* __restrict__ keys, const int* __restrict__ slots, int n, int world, OldCsr oc, const uint32_t* __restrict__ oflags, int* rowCnt) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  int c = 0;
  if (k < n && oc.valid) {
    const uint32_t key = keys[k];
    if (key >= 1u && key <= (uint32_t)world && slots[k] < oc.nRows) { const int s = slots[k]; for (int e = oc.rowStart[s]; e < oc.rowStart[s + 1]; e++) { const uint32_t f = oflags[oc.rowEntry[e] >> 1]; if ((f & CF_REAL) && !(f & CF_DEAD)) c++; } }
  }
  rowCnt[k] = c;
}
// one thread: first index of every destination key (empty keys inherit the next one's), and what goes to every rank: (spheres, history rows)
__global__ void k_dom_send_counts(int* start /*[world+3], pre-filled with n by k_dom_starts*/, int n, int world, const int* __restrict__ rowOff, int* perDest /*[2*world]*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  start[world + 2] = n;
  for (int k = world + 1; k >= 0; k--) if (start[k] == n) start[k] = start[k + 1];
  for (int r = 0; r < world; r++) { const int a = start[1 + r], b = start[2 + r]; perDest[2 * r] = b - a; perDest[2 * r + 1] = rowOff[b] - rowOff[a]; }
}

__device__ __forceinline__ void crec_from(const Contacts& O, long long co, int idA, int idB, int law, int rolling, int periodic, CRec& r) {
  r.idA = idA; r.idB = idB; r.flags = O.flags[co]; r.stepCreated = O.stepCreated[co]; r.pad = 0;
  r.cd[0] = periodic ? O.cdx[co] : 0; r.cd[1] = periodic ? O.cdy[co] : 0; r.cd[2] = periodic ? O.cdz[co] : 0;
  const double4 r0 = O.rA[co], r1 = O.rB[co], r3 = O.rD[co], p = O.ph[co];
  r.t[0] = r0.x; r.t[1] = r0.y; r.t[2] = r0.z; r.t[3] = r1.x; r.t[4] = r1.y; r.t[5] = r1.z;
  r.cp = O.rC[co]; r.fl[0] = r3.w; r.fl[1] = r0.w; r.fl[2] = r1.w;
  r.ph[0] = p.x; r.ph[1] = p.y; r.ph[2] = p.z;
  r.ph[3] = law == 1 ? p.w : 0.; r.ph[4] = law == 1 ? O.ph4[co] : 0.;
  double4 tq = make_double4(0, 0, 0, 0); if (rolling) tq = O.tq[co];
  r.tq[0] = tq.x; r.tq[1] = tq.y; r.tq[2] = tq.z;
  r.Fg = make_double4(r3.x, r3.y, r3.z, 0.);
}
__global__ void __launch_bounds__(128) k_dom_pack_rows(const int* __restrict__ slots, int first, int cnt, OldCsr oc, Contacts O, const int32_t* __restrict__ sid,
                                                      const Wall* __restrict__ walls, const Facet* __restrict__ facets, const int* __restrict__ rowOff, int law, int rolling,
                                                      int periodic, CRec* out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt || !oc.valid) return;
  const int s = slots[first + k];
  if (s >= oc.nRows) return;      // appended since the rows were built (wdem_add_spheres): no history
  int w = rowOff[k];
  for (int e = oc.rowStart[s]; e < oc.rowStart[s + 1]; e++) {
    const long long co = oc.rowEntry[e] >> 1;
    const uint32_t f = O.flags[co];
    if (!(f & CF_REAL) || (f & CF_DEAD)) continue;
    const int2 pr = O.pair[co]; const int kind = cf_kind(f);
    const int idA = kind == KIND_SS ? sid[pr.x] : (kind == KIND_WS ? walls[pr.x].id : facets[pr.x].id);
    crec_from(O, co, idA, sid[pr.y], law, rolling, periodic, out[w]);
    out[w].pad = pr.x;   // wall / facet index for boundary rows
    w++;
  }
}
// received rows are appended behind the previous table (so that k_carry can index them) and keyed for the lookup
__global__ void __launch_bounds__(256) k_dom_unpack_rows(Contacts O, long long base, int cnt, const CRec* __restrict__ in, int law, int rolling, int periodic,
                                                        unsigned long long* keys, int* vals, int uploaded) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const CRec r = in[k];
  const long long co = base + k;
  // migrated rows are never orphan candidates (the sender keeps or drops them by its own rules); rows imported by wdem_upload_contacts carry the
  // slots of their particles in the current generation and become orphans if no candidate claims them (real although the AABBs are apart)
  O.pair[co] = uploaded ? make_int2(r.pad, r.pairY) : make_int2(r.pad, -1);
  O.flags[co] = (r.flags & (CF_KIND_MASK | CF_REAL)) | (uploaded ? 0u : (uint32_t)CF_CLAIMED);
  O.stepCreated[co] = r.stepCreated;
  if (periodic) { O.cdx[co] = r.cd[0]; O.cdy[co] = r.cd[1]; O.cdz[co] = r.cd[2]; }
  O.rA[co] = make_double4(r.t[0], r.t[1], r.t[2], r.fl[1]); O.rB[co] = make_double4(r.t[3], r.t[4], r.t[5], r.fl[2]);
  O.rC[co] = r.cp; O.rD[co] = make_double4(r.Fg.x, r.Fg.y, r.Fg.z, r.fl[0]);
  O.ph[co] = make_double4(r.ph[0], r.ph[1], r.ph[2], law == 1 ? r.ph[3] : 0.);
  if (law == 1) O.ph4[co] = r.ph[4];
  if (rolling) O.tq[co] = make_double4(r.tq[0], r.tq[1], r.tq[2], 0.);
  keys[k] = ((unsigned long long)(uint32_t)r.idA << 32) | (uint32_t)r.idB;
  vals[k] = (int)co;
}

// ghost export: bit r of mask[i] set if the enlarged box of owned sphere i can overlap the enlarged box of a sphere owned by rank r
// Full run without migration (the usual case in a bed at rest), one pass over the current generation IN PLACE: counts the owned spheres that left the slab
// (out[0]; any rank > 0 => everybody takes the regrouping path) and, for the owned spheres, the ghost-export mask and per-destination counts (out[1+r]).
// Ghosts of the previous generation get mask 0.
__global__ void __launch_bounds__(256) k_dom_fast_scan(Spheres S, int n, DomainU D, uint32_t* mask, int* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0, emig = 0;
  if (i < n) {
    if (!(S.flags[i] & F_GHOST)) {
      const double4 p = S.posr[i];
      const double x = slabCoord(D, p), h = slabHalf(D, p) + D.reach;
      emig = ownerOf(D, x) != D.rank;
      for (int r = 0; r < D.world; r++) {
        if (r == D.rank) continue;
        const double lo = D.bounds[r], hi = D.bounds[r + 1];
        bool hit;
        if (!D.periodic) hit = (x + h >= lo) && (x - h <= hi);
        else {
          if (x >= lo && x < hi) hit = true;
          else { const double d1 = wrap01(lo - x, D.dim), d2 = wrap01(x - hi, D.dim); hit = fmin(d1, d2) <= h; }
        }
        if (hit) m |= 1u << r;
      }
    }
    mask[i] = m;
  }
  const unsigned e = __reduce_add_sync(0xffffffffu, emig);
  if ((threadIdx.x & 31) == 0 && e) atomicAdd(&out[0], (int)e);
  for (int r = 0; r < D.world; r++) {
    const unsigned c = __reduce_add_sync(0xffffffffu, (m >> r) & 1u);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&out[1 + r], (int)c);
  }
}
// re-balancing of the cuts at a full run: histogram of the slab coordinate of the owned spheres over [lo, lo + nb/inv) (coordinates outside fall into the end bins)
#define WD_REBAL_BINS 2048
__global__ void __launch_bounds__(256) k_dom_coord_hist(Spheres S, int n, DomainU D, double lo, double inv, int* hist /*[WD_REBAL_BINS]*/) {
  __shared__ int sh[WD_REBAL_BINS];
  for (int k = threadIdx.x; k < WD_REBAL_BINS; k += blockDim.x) sh[k] = 0;
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(S.flags[i] & F_GHOST)) {
    const double x = (slabCoord(D, S.posr[i]) - lo) * inv;
    const int b = x < 0. ? 0 : (x >= (double)(WD_REBAL_BINS - 1) ? WD_REBAL_BINS - 1 : (int)x);
    atomicAdd(&sh[b], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < WD_REBAL_BINS; k += blockDim.x) if (sh[k]) atomicAdd(&hist[k], sh[k]);
}
// wdem_add_spheres / wdem_remove_spheres in decomposed runs (every rank gets the same list).
// is any of these ids held here (owned sphere or ghost copy)?
__global__ void __launch_bounds__(256) k_dom_ids_in_use(const int32_t* __restrict__ ids, int m, const int* __restrict__ id2slot, long long cap, int n, int* flag) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int id = ids[k];
  if (id >= 0 && id < cap) { const int s = id2slot[id]; if (s >= 0 && s < n) atomicMax(flag, id + 1); }
}
// removal: whatever this subdomain holds of the listed spheres stops being owned (the next full run drops it like a ghost of the previous generation); keep[slot]=0 for k_kill_rows
__global__ void __launch_bounds__(256) k_dom_mark_removed(const int32_t* __restrict__ ids, int m, const int* __restrict__ id2slot, long long cap, int n, uint32_t* flags, int* keep, int* nOwnedGone) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int id = ids[k];
  if (id < 0 || id >= cap) return;
  const int s = id2slot[id];
  if (s < 0 || s >= n) return;
  keep[s] = 0;
  if (!(flags[s] & F_GHOST)) atomicAdd(nOwnedGone, 1);
  flags[s] |= F_GHOST;
}
// wdem_step_e2e_owned: slots of the ids the host sent (its order); an id this subdomain does not own counts as a miss and maps to slot 0 of a ghost-safe no-op
__global__ void __launch_bounds__(256) k_owned_slots(const int32_t* __restrict__ ids, int m, const int* __restrict__ id2slot, long long cap, int n, const uint32_t* __restrict__ flags,
                                                     int* __restrict__ slots, int* miss) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int id = ids[k]; int s = -1;
  if (id >= 0 && id < cap) { s = id2slot[id]; if (s < 0 || s >= n || (flags[s] & F_GHOST)) s = -1; }
  if (s < 0) atomicAdd(miss, 1);
  slots[k] = s;
}
// appended spheres are addressable by id at once (a later add / remove in the same breath must see them)
__global__ void __launch_bounds__(256) k_dom_map_appended(const int32_t* __restrict__ sid, int first, int cnt, int* id2slot) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < cnt) id2slot[sid[first + k]] = first + k;
}
__global__ void __launch_bounds__(256) k_dom_forget_ids(const int32_t* __restrict__ ids, int m, int* id2slot, long long cap) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < m && ids[k] >= 0 && ids[k] < cap) id2slot[ids[k]] = -1;
}
// the ghosts of the previous generation leave in place: their slots are cut off by the full run that follows
__global__ void __launch_bounds__(256) k_dom_mark_dead(uint32_t* flags, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t f = flags[i];
  flags[i] = (f & F_GHOST) ? (f | F_DEADSLOT) : (f & ~F_ARRIVED);
}
__global__ void __launch_bounds__(256) k_dom_export_mask(Spheres S, int nOwn, DomainU D, uint32_t* mask, int* perRank) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t m = 0;
  if (i < nOwn) {
    const double4 p = S.posr[i];
    const double x = slabCoord(D, p), h = slabHalf(D, p) + D.reach;
    for (int r = 0; r < D.world; r++) {
      if (r == D.rank) continue;
      const double lo = D.bounds[r], hi = D.bounds[r + 1];
      bool hit;
      if (!D.periodic) hit = (x + h >= lo) && (x - h <= hi);
      else {
        if (x >= lo && x < hi) hit = true;
        else { const double d1 = wrap01(lo - x, D.dim), d2 = wrap01(x - hi, D.dim); hit = fmin(d1, d2) <= h; }
      }
      if (hit) m |= 1u << r;
    }
    mask[i] = m;
  }
  for (int r = 0; r < D.world; r++) {
    const unsigned c = __reduce_add_sync(0xffffffffu, (m >> r) & 1u);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&perRank[r], (int)c);
  }
}
__global__ void __launch_bounds__(256) k_dom_export_flag(const uint32_t* __restrict__ mask, int nOwn, int r, int* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nOwn) flag[i] = (mask[i] >> r) & 1u;
}
__global__ void __launch_bounds__(256) k_dom_pack_ghosts(Spheres S, int nOwn, const uint32_t* __restrict__ mask, int r, const int* __restrict__ off, GRec* out, int32_t* exportId) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nOwn || !((mask[i] >> r) & 1u)) return;
  GRec g; g.posr = S.posr[i]; g.velm = S.velm[i]; g.angi = S.angi[i]; g.flags = (S.flags[i] & ~F_ARRIVED) | F_GHOST; g.mask = S.mask[i]; g.mat = S.mat[i]; g.id = S.id[i];
  out[off[i]] = g; exportId[off[i]] = g.id;
}
// all destinations in one pass: a sphere exported to rank r takes the next free place of r's segment (atomic cursor).  The order inside a segment is arbitrary
// but the SAME on both sides: exportId here, importId on the receiver are filled from the same records.
struct GhostSeg { int base[WDEM_MAX_RANKS]; };
struct GhostDst { GRec* out[WDEM_MAX_RANKS]; };   // where destination r's records go: a segment of my send buffer (NCCL path) or of r's own staging area (peer memory)
__global__ void __launch_bounds__(256) k_dom_pack_ghosts_all(Spheres S, int nOwn, const uint32_t* __restrict__ mask, int world, GhostSeg seg, int* __restrict__ cursor, GhostDst dst, int32_t* exportId) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nOwn) return;
  uint32_t m = mask[i];
  if (!m) return;
  GRec g; g.posr = S.posr[i]; g.velm = S.velm[i]; g.angi = S.angi[i]; g.flags = (S.flags[i] & ~F_ARRIVED) | F_GHOST; g.mask = S.mask[i]; g.mat = S.mat[i]; g.id = S.id[i];
  while (m) {
    const int r = __ffs(m) - 1; m &= m - 1;
    const int k = atomicAdd(&cursor[r], 1);
    dst.out[r][k] = g; exportId[seg.base[r] + k] = g.id;
  }
}
__global__ void __launch_bounds__(256) k_dom_unpack_ghosts(Spheres S, int first, int cnt, const GRec* __restrict__ in, int32_t* importId) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= cnt) return;
  const int i = first + k; const GRec g = in[k];
  S.posr[i] = g.posr; S.velm[i] = g.velm; S.angi[i] = g.angi; S.ori[i] = make_double4(0, 0, 0, 1);
  S.fx[i] = S.fy[i] = S.fz[i] = S.tx[i] = S.ty[i] = S.tz[i] = 0.;
  S.lx[i] = g.posr.x; S.ly[i] = g.posr.y; S.lz[i] = g.posr.z;
  S.flags[i] = g.flags | F_GHOST; S.mask[i] = g.mask; S.mat[i] = g.mat; S.id[i] = g.id; S.up[i] = -1;
  importId[k] = g.id;
}
// slot of each sphere of the new (pre-sort) generation in the arrays the previous contact table / CSR refer to (-1: was not here)
__global__ void __launch_bounds__(256) k_dom_orig_slot(const int32_t* __restrict__ id, int n, const int* __restrict__ id2slot, int* origSlot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) origSlot[i] = id2slot[id[i]];
}
__global__ void __launch_bounds__(256) k_dom_clear_id2slot(const int32_t* __restrict__ id, int n, int* id2slot) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id2slot[id[i]] = -1;
}
__global__ void __launch_bounds__(256) k_dom_compose(const int* __restrict__ origSlot, const int* __restrict__ perm, int n, int* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = origSlot[perm[i]];
}
__global__ void __launch_bounds__(256) k_dom_ids_to_slots(const int32_t* __restrict__ ids, int n, const int* __restrict__ id2slot, int* slots) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) slots[i] = id2slot[ids[i]];
}

// per-step ghost refresh: 9 doubles per exported sphere
__global__ void __launch_bounds__(256) k_halo_pack(Spheres S, const int* __restrict__ slots, int n, double* buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = slots[k];
  const double4 p = S.posr[i], v = S.velm[i], w = S.angi[i];
  double* o = buf + 9ll * k;
  o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = v.x; o[4] = v.y; o[5] = v.z; o[6] = w.x; o[7] = w.y; o[8] = w.z;
}
__global__ void __launch_bounds__(256) k_halo_unpack(Spheres S, const int* __restrict__ slots, int n, const double* __restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int i = slots[k];
  const double* o = buf + 9ll * k;
  double4 p = S.posr[i], v = S.velm[i], w = S.angi[i];
  p.x = o[0]; p.y = o[1]; p.z = o[2]; v.x = o[3]; v.y = o[4]; v.z = o[5]; w.x = o[6]; w.y = o[7]; w.z = o[8];
  S.posr[i] = p; S.velm[i] = v; S.angi[i] = w;
}

// ------------------------------------------------------------------------------------------------------------------------------------
// Per-step exchange over NVLink peer memory (one process per GPU; mailboxes mapped with CUDA IPC): instead of ncclAllReduce + ncclSend/ncclRecv
// every owner stores the pos/vel/angVel of its exported spheres straight into the neighbours' halo buffers and then publishes ONE 64-bit word
// per peer, (step stamp | device error | Verlet-dirty), with release semantics at system scope; the receiver spins on the words of all ranks
// (acquire), which gives it the global dirty/error maximum and the guarantee that the halo payload written before the word has landed.
// Buffers and words are double-buffered by step parity: a rank can run at most one step ahead of a peer (it needs the peer's word of step k to
// finish step k), so the slot of step k+2 is never written before the peer has consumed step k.
// ------------------------------------------------------------------------------------------------------------------------------------
struct PeerU {
  int rank, world;
  unsigned long long* words[WDEM_MAX_RANKS];   // peer r's word array [2][WDEM_MAX_RANKS] (mine included)
  double* halo[WDEM_MAX_RANKS];                // peer r's halo buffers: 2 x haloCap doubles
  long long haloCap;                           // doubles per parity buffer
  int sendBase[WDEM_MAX_RANKS + 1];            // my export list: entries [sendBase[r], sendBase[r+1]) go to rank r
  long long dstOff[WDEM_MAX_RANKS];            // where my segment starts in rank r's halo buffer (in spheres)
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

__global__ void __launch_bounds__(256) k_halo_push(Spheres S, const int* __restrict__ slots, int n, PeerU P, int parity) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int r = 0;
  while (k >= P.sendBase[r + 1]) r++;
  const int i = slots[k];
  const double4 p = S.posr[i], v = S.velm[i], w = S.angi[i];
  double* o = P.halo[r] + (long long)parity * P.haloCap + 9ll * (P.dstOff[r] + (k - P.sendBase[r]));
  o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = v.x; o[4] = v.y; o[5] = v.z; o[6] = w.x; o[7] = w.y; o[8] = w.z;
}
// one CTA, one thread per rank: publish my word to every peer (after the halo payload: previous kernel on the stream + system fence), then wait for
// theirs.  cnt->dirty / cnt->error become the maxima over all ranks.  A peer that never answers (it failed on the host) ends the wait after
// `timeoutCycles` with ERR_PEER_TIMEOUT instead of hanging the GPU.
__global__ void k_peer_sync(PeerU P, Counters* cnt, int* flag /*this step's Verlet-dirty slot*/, unsigned long long stamp, int parity, long long timeoutCycles) {
  __shared__ int sDirty[WDEM_MAX_RANKS], sErr[WDEM_MAX_RANKS];
  const int r = threadIdx.x;
  const unsigned long long mine = (stamp << 16) | ((unsigned long long)(cnt->error & 0x7fff) << 1) | (unsigned long long)(*flag ? 1 : 0);
  int d = 0, e = 0;
  if (r < P.world) {
    __threadfence_system();
    st_release_sys(P.words[r] + parity * WDEM_MAX_RANKS + P.rank, mine);
    const unsigned long long* w = P.words[P.rank] + parity * WDEM_MAX_RANKS + r;
    const long long t0 = clock64();
    unsigned long long v = ld_acquire_sys(w);
    while ((v >> 16) != stamp) {
      if (clock64() - t0 > timeoutCycles) { v = (stamp << 16) | ((unsigned long long)ERR_PEER_TIMEOUT << 1) | 1ull; break; }
      __nanosleep(100);
      v = ld_acquire_sys(w);
    }
    d = (int)(v & 1ull); e = (int)((v >> 1) & 0x7fff);
  }
  sDirty[r] = d; sErr[r] = e;
  __syncthreads();
  if (r == 0) {
    int dm = 0, em = 0;
    for (int q = 0; q < P.world; q++) { dm = max(dm, sDirty[q]); em = max(em, sErr[q]); }
    *flag = dm;
    if (em && !cnt->error) { cnt->error = em; cnt->errA = -1; cnt->errB = -1; }
    __threadfence();
  }
}

// contacts counted once over all ranks: by the owner of the sphere that is pA (sphere-sphere) / of the sphere (boundary contacts)
__global__ void __launch_bounds__(256) k_count_contacts_owned(Contacts C, long long n, const uint32_t* __restrict__ sflags, unsigned long long* out) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned r = 0, a = 0;
  if (c < n) {
    const uint32_t f = C.flags[c]; const int2 pr = C.pair[c];
    const int s = cf_kind(f) == KIND_SS ? pr.x : pr.y;
    if (!(sflags[s] & F_GHOST)) { r = (f & CF_REAL) ? 1 : 0; a = (f & CF_DEAD) ? 0 : 1; }
  }
  r = __reduce_add_sync(0xffffffffu, r); a = __reduce_add_sync(0xffffffffu, a);
  if ((threadIdx.x & 31) == 0) { if (r) atomicAdd(&out[0], (unsigned long long)r); if (a) atomicAdd(&out[1], (unsigned long long)a); }
}

// owned spheres, compacted, with their ids (downloads in decomposed runs)
__global__ void __launch_bounds__(256) k_gather_owned(Spheres S, int n, const int* __restrict__ off, int32_t* ids, double* pos, double* ori, double* vel, double* angVel, double* force,
                                                     double* torque) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (S.flags[i] & F_GHOST)) return;
  const long long u = off[i];
  ids[u] = S.id[i];
  if (pos) { double4 p = S.posr[i]; pos[3 * u] = p.x; pos[3 * u + 1] = p.y; pos[3 * u + 2] = p.z; }
  if (ori) { double4 q = S.ori[i]; ori[4 * u] = q.x; ori[4 * u + 1] = q.y; ori[4 * u + 2] = q.z; ori[4 * u + 3] = q.w; }
  if (vel) { double4 v = S.velm[i]; vel[3 * u] = v.x; vel[3 * u + 1] = v.y; vel[3 * u + 2] = v.z; }
  if (angVel) { double4 w = S.angi[i]; angVel[3 * u] = w.x; angVel[3 * u + 1] = w.y; angVel[3 * u + 2] = w.z; }
  if (force) { force[3 * u] = S.fx[i]; force[3 * u + 1] = S.fy[i]; force[3 * u + 2] = S.fz[i]; }
  if (torque) { torque[3 * u] = S.tx[i]; torque[3 * u + 1] = S.ty[i]; torque[3 * u + 2] = S.tz[i]; }
}
// host-supplied state of owned spheres addressed by Particle.id (end-to-end path of decomposed runs)
__global__ void __launch_bounds__(256) k_scatter_owned(Spheres S, const int* __restrict__ id2slot, const int32_t* __restrict__ ids, long long cnt, const double* pos,
                                                      const double* vel, const double* angVel, int* missing) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int i = id2slot[ids[t]];
  if (i < 0 || (S.flags[i] & F_GHOST)) { atomicAdd(missing, 1); return; }
  if (pos) { double4 p = S.posr[i]; p.x = pos[3 * t]; p.y = pos[3 * t + 1]; p.z = pos[3 * t + 2]; S.posr[i] = p; }
  if (vel) { double4 v = S.velm[i]; v.x = vel[3 * t]; v.y = vel[3 * t + 1]; v.z = vel[3 * t + 2]; S.velm[i] = v; }
  if (angVel) { double4 w = S.angi[i]; w.x = angVel[3 * t]; w.y = angVel[3 * t + 1]; w.z = angVel[3 * t + 2]; S.angi[i] = w; }
}
__global__ void __launch_bounds__(256) k_owned_flag(const uint32_t* __restrict__ sflags, int n, int* flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (sflags[i] & F_GHOST) ? 0 : 1;
}

}  // namespace wd
