// Host-side check vehicle for the batched-affine bucket accumulation (blaze_b200/csrc/msm_ba2.cuh): the per-segment
// routine the CUDA kernel runs per lane, compiled with g++ against the carry-flag emulation of bz_common.cuh and driven
// segment by segment with the kernel's own scratch layout.  Test-only (tests/test_ba_on_host.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../blaze_b200/csrc/msm_ba2.cuh"

using namespace bz;

template <class C>
static int run(const uint8_t* pts, uint32_t npoints, const uint32_t* sorted, uint32_t total, const uint32_t* goff, uint32_t ngoff,
               uint32_t L, int rounds, uint32_t cap, uint8_t* out) {
  typedef ba2<C> B;
  typedef ff<typename C::Fq> F;
  typedef ec<C> G;
  constexpr int N = C::Fq::N;
  std::vector<AffineT<C>> table(npoints);
  for (uint32_t i = 0; i < npoints; i++) {
    Fe<typename C::Fq> x = F::zero(), y = F::zero();
    memcpy(x.v, pts + (size_t)i * 2 * C::FQ_BYTES, C::FQ_BYTES);
    memcpy(y.v, pts + (size_t)i * 2 * C::FQ_BYTES + C::FQ_BYTES, C::FQ_BYTES);
    x = F::to_mont(x);
    y = F::to_mont(y);
    memcpy(table[i].x, x.v, 4 * N);
    memcpy(table[i].y, y.v, 4 * N);
  }
  const uint64_t nseg = ((uint64_t)total + L - 1) / L;
  std::vector<XyzzM<C>> buckets(ngoff), part_pt(2 * nseg + 2);
  memset(buckets.data(), 0, buckets.size() * sizeof(XyzzM<C>));
  std::vector<uint32_t> part_id(2 * nseg + 2, 0xffffffffu);
  const size_t plane = (size_t)cap * B::NQ * 32;
  std::vector<typename B::Q> scratch(3 * plane);
  for (uint64_t t = 0; t < nseg; t++) {
    const uint32_t lane = (uint32_t)(t & 31);
    typename B::Ctx c;
    c.table = table.data();
    c.sorted = sorted;
    c.goff = goff;
    c.ngoff = ngoff;
    c.cap = cap;
    c.sx = scratch.data() + lane;
    c.sy = c.sx + plane;
    c.sp = c.sy + plane;
    c.s = (uint32_t)(t * L);
    c.e = (uint32_t)(t * L + L < total ? t * L + L : total);
    uint32_t lo = 0, hi = ngoff;
    while (hi - lo > 1) {
      uint32_t mid = lo + ((hi - lo) >> 1);
      if (goff[mid] <= c.s) lo = mid; else hi = mid;
    }
    c.g_first = lo;
    B::segment(c, t, rounds, buckets.data(), part_id.data(), part_pt.data());
  }
  auto load = [](const XyzzM<C>& m) {
    XYZZ<C> p;
    memcpy(p.X.v, m.X, 4 * N); memcpy(p.Y.v, m.Y, 4 * N); memcpy(p.ZZ.v, m.ZZ, 4 * N); memcpy(p.ZZZ.v, m.ZZZ, 4 * N);
    return p;
  };
  std::vector<XYZZ<C>> sum(ngoff);
  for (uint32_t g = 0; g < ngoff; g++) sum[g] = load(buckets[g]);
  for (uint64_t k = 0; k < 2 * nseg; k++) {
    if (part_id[k] == 0xffffffffu) continue;
    if (part_id[k] >= ngoff) return -2;
    XYZZ<C> p = load(part_pt[k]);
    G::add(sum[part_id[k]], p);
  }
  memset(out, 0, (size_t)ngoff * 2 * C::FQ_BYTES);
  for (uint32_t g = 0; g < ngoff; g++) {
    Affine<C> a;
    if (!G::to_affine(sum[g], a)) continue;
    Fe<typename C::Fq> x = F::from_mont(a.x), y = F::from_mont(a.y);
    memcpy(out + (size_t)g * 2 * C::FQ_BYTES, x.v, C::FQ_BYTES);
    memcpy(out + (size_t)g * 2 * C::FQ_BYTES + C::FQ_BYTES, y.v, C::FQ_BYTES);
  }
  return 0;
}

extern "C" int hc_ba_accumulate(int curve, const uint8_t* pts, uint32_t npoints, const uint32_t* sorted, uint32_t total,
                                const uint32_t* goff, uint32_t ngoff, uint32_t L, int rounds, uint32_t cap, uint8_t* out) {
  switch (curve) {
    case 0: return run<Bls12_377>(pts, npoints, sorted, total, goff, ngoff, L, rounds, cap, out);
    case 1: return run<Bn254>(pts, npoints, sorted, total, goff, ngoff, L, rounds, cap, out);
    case 2: return run<Bls12_381>(pts, npoints, sorted, total, goff, ngoff, L, rounds, cap, out);
  }
  return -1;
}
