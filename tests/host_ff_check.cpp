// Host-side check vehicle for the DEVICE field / curve code (blaze_b200/csrc/ff.cuh, ec.cuh).
// Compiled with g++: bz_common.cuh then emulates the PTX carry-chain instructions with a
// thread-local flag, so the exact limb schedules that run on the GPU are exercised here
// bit-for-bit against the oracle (tests/test_device_math_on_host.py).  Test-only; the product
// library never runs field code on the host.
#include <cstdint>
#include <cstring>

#include "../blaze_b200/csrc/ec.cuh"

using namespace bz;

template <class F>
static Fe<F> load(const uint8_t* b, int nbytes) {
  Fe<F> r = ff<F>::zero();
  memcpy(r.v, b, nbytes);
  return ff<F>::to_mont(r);
}
template <class F>
static void store(uint8_t* b, const Fe<F>& a, int nbytes) {
  Fe<F> r = ff<F>::from_mont(a);
  memcpy(b, r.v, nbytes);
}

template <class F>
static int field_op(int op, const uint8_t* a, const uint8_t* b, uint8_t* out, int nbytes) {
  Fe<F> x = load<F>(a, nbytes), y = load<F>(b, nbytes), r;
  switch (op) {
    case 0: r = ff<F>::mul(x, y); break;
    case 1: r = ff<F>::add(x, y); break;
    case 2: r = ff<F>::sub(x, y); break;
    case 3: r = ff<F>::sqr(x); break;
    case 4: r = ff<F>::inv(x); break;
    case 5: r = ff<F>::neg(x); break;
    case 6: r = ff<F>::dbl(x); break;
    case 7: r = ff<F>::mul_sub2(x, y, ff<F>::add(x, y), ff<F>::sub(x, y)); break;   // xy - (x+y)(x-y)
    case 8: r = ff<F>::mul2(x, y, ff<F>::add(x, y), ff<F>::sub(x, y)); break;
    case 11: r = ff<F>::inv_gcd(x); break;   // division-step inverse
    case 12: r = ff<F>::inv_fermat(x); break;
    case 9: r = ff<F>::mul_kara(x, y); break;   // Karatsuba product + reduction-only Montgomery
    case 10:
      if constexpr (F::BITS + 2 <= 32 * F::N) r = ff<F>::mul2_kara(x, y, ff<F>::add(x, y), ff<F>::sub(x, y));
      else r = ff<F>::add(ff<F>::mul_kara(x, y), ff<F>::mul_kara(ff<F>::add(x, y), ff<F>::sub(x, y)));
      break;
    case 13: r = ff<F>::sqr_split(x); break;              // squaring product + reduction-only Montgomery (BZ_SPLIT_MUL)
    case 14: r = ff<F>::mul_split(x, y); break;           // redc(prod): the split-call product
    case 15:
      if constexpr (F::BITS + 2 <= 32 * F::N) r = ff<F>::mul2_split(x, y, ff<F>::add(x, y), ff<F>::sub(x, y));
      else r = ff<F>::add(ff<F>::mul_split(x, y), ff<F>::mul_split(ff<F>::add(x, y), ff<F>::sub(x, y)));
      break;
    default: return -1;
  }
  store<F>(out, r, nbytes);
  return 0;
}

template <class C>
static Affine<C> load_pt(const uint8_t* b) {
  Affine<C> a;
  a.x = load<typename C::Fq>(b, C::FQ_BYTES);
  a.y = load<typename C::Fq>(b + C::FQ_BYTES, C::FQ_BYTES);
  return a;
}

// ops: 0: k1*P + k2*Q using madd/add/dbl/mul_small ;  returns 1 for infinity
template <class C>
static int curve_op(int op, const uint8_t* p, const uint8_t* q, uint32_t k1, uint32_t k2, uint8_t* out) {
  typedef ec<C> E;
  Affine<C> a = load_pt<C>(p), b = load_pt<C>(q);
  XYZZ<C> acc = E::infinity();
  if (op == 0) {          // repeated mixed adds: k1 times P then k2 times Q (exercises doubling branch)
    for (uint32_t i = 0; i < k1; i++) E::madd(acc, a);
    for (uint32_t i = 0; i < k2; i++) E::madd(acc, b);
  } else if (op == 1) {   // mul_small + full add
    XYZZ<C> x = E::mul_small(E::from_affine(a), k1);
    XYZZ<C> y = E::mul_small(E::from_affine(b), k2);
    acc = x;
    E::add(acc, y);
  } else if (op == 2) {   // P + (-Q) via neg, then full add of itself (dbl branch)
    E::madd(acc, a);
    E::madd(acc, E::neg(b));
    XYZZ<C> t = acc;
    E::add(acc, t);
  } else if (op == 3) {   // batched-affine building blocks: classify, invert the denominator, finish
    typename E::E den;
    int kind = E::ba_classify(a, b, den);
    Affine<C> r3 = E::ba_finish(kind, a, b, ff<typename C::Fq>::inv(den));
    memset(out, 0, 2 * C::FQ_BYTES);
    if (E::is_identity(r3)) return 1;
    store<typename C::Fq>(out, r3.x, C::FQ_BYTES);
    store<typename C::Fq>(out + C::FQ_BYTES, r3.y, C::FQ_BYTES);
    return 0;
  } else {
    return -1;
  }
  Affine<C> r;
  memset(out, 0, 2 * C::FQ_BYTES);
  if (!E::to_affine(acc, r)) return 1;
  store<typename C::Fq>(out, r.x, C::FQ_BYTES);
  store<typename C::Fq>(out + C::FQ_BYTES, r.y, C::FQ_BYTES);
  return 0;
}

extern "C" {
// field ids: 0 Fq377, 1 Fq254, 2 Fq381, 10 Fr377, 11 Fr254, 12 Fr381
int hc_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  switch (field) {
    case 0: return field_op<Fq377>(op, a, b, out, 48);
    case 1: return field_op<Fq254>(op, a, b, out, 32);
    case 2: return field_op<Fq381>(op, a, b, out, 48);
    case 10: return field_op<Fr377>(op, a, b, out, 32);
    case 11: return field_op<Fr254>(op, a, b, out, 32);
    case 12: return field_op<Fr381>(op, a, b, out, 32);
  }
  return -1;
}
int hc_curve_op(int curve, int op, const uint8_t* p, const uint8_t* q, uint32_t k1, uint32_t k2, uint8_t* out) {
  switch (curve) {
    case 0: return curve_op<Bls12_377>(op, p, q, k1, k2, out);
    case 1: return curve_op<Bn254>(op, p, q, k1, k2, out);
    case 2: return curve_op<Bls12_381>(op, p, q, k1, k2, out);
  }
  return -1;
}
}
