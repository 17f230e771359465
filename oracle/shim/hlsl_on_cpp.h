// TEST INFRASTRUCTURE — not product code.
//
// A small "HLSL vector language on C++17" shim.  It exists for one purpose: to let g++ compile
// the reference's own A_GPU+A_HLSL source (ffx-fsr/ffx_a.h + ffx-fsr/ffx_fsr1.h, read from
// /root/reference at build time, never copied into this repo) on the host, so that the
// reference's per-pixel functions FsrEasuF/FsrRcasF (and, with FSR1_REF_HALF, FsrEasuH/FsrRcasH)
// can be executed here and pin the oracle.  See oracle/build_ref.sh and oracle/ref_driver.cpp.
//
// What HLSL semantics are mirrored (they matter for parity):
//   * min/max are IEEE minNum/maxNum (return the non-NaN operand)  -> fminf/fmaxf
//   * saturate(NaN) == 0                                           -> fminf(fmaxf(x,0),1)
//   * rcp(x) == 1/x, rsqrt(x) == 1/sqrt(x) (IEEE division)
//   * asuint/asfloat are bit casts; f32tof16/f16tof32 are IEEE round-to-nearest-even
//   * min16float is modelled as _Float16 with per-operation rounding (-fexcess-precision=16)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

typedef unsigned int uint;
#ifdef FSR1_REF_HALF
typedef _Float16 min16float;
typedef uint16_t min16uint;
typedef int16_t min16int;
#endif

namespace hlsl {

template <class T, int N> struct vec;

// ---- swizzle proxy: lives inside the vec's union, aliases its storage ------------------------
template <class T, int N, int... I> struct swz {
  T d[N];
  static constexpr int M = sizeof...(I);
  operator vec<T, M>() const {
    vec<T, M> r;
    const int idx[M] = {I...};
    for (int i = 0; i < M; i++) r.d[i] = d[idx[i]];
    return r;
  }
  swz& operator=(const vec<T, M>& v) {
    const int idx[M] = {I...};
    for (int i = 0; i < M; i++) d[idx[i]] = v.d[i];
    return *this;
  }
  swz& operator=(const swz& o) { return *this = (vec<T, M>)o; }
  template <int... J> swz& operator=(const swz<T, N, J...>& o) { return *this = (vec<T, M>)o; }
};

template <class T> struct vec<T, 2> {
  union {
    T d[2];
    struct { T x, y; };
    struct { T r, g; };
    swz<T, 2, 0, 0> xx; swz<T, 2, 1, 1> yy; swz<T, 2, 0, 1> xy; swz<T, 2, 1, 0> yx;
    swz<T, 2, 0, 0, 0> xxx; swz<T, 2, 1, 1, 1> yyy;
  };
  vec() : d{} {}
  vec(T a, T b) : d{a, b} {}
  template <class A, class B, class = std::enable_if_t<std::is_arithmetic_v<A> || std::is_same_v<A, T>>,
            class = std::enable_if_t<std::is_arithmetic_v<B> || std::is_same_v<B, T>>>
  vec(A a, B b) : d{(T)a, (T)b} {}
  template <class S, class = std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, T>>>
  vec(S s) : d{(T)s, (T)s} {}
  template <class U> explicit vec(const vec<U, 2>& o) : d{(T)o.d[0], (T)o.d[1]} {}
  template <class U, int NN, int A, int B> explicit vec(const swz<U, NN, A, B>& o) {
    vec<U, 2> t = o; d[0] = (T)t.d[0]; d[1] = (T)t.d[1];
  }
  vec(const vec& o) { d[0] = o.d[0]; d[1] = o.d[1]; }
  vec& operator=(const vec& o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
};

template <class T> struct vec<T, 3> {
  union {
    T d[3];
    struct { T x, y, z; };
    struct { T r, g, b; };
    swz<T, 3, 0, 1> xy; swz<T, 3, 1, 2> yz; swz<T, 3, 0, 1> rg;
    swz<T, 3, 0, 0> xx; swz<T, 3, 1, 1> yy; swz<T, 3, 2, 2> zz;
    swz<T, 3, 0, 0, 0> xxx; swz<T, 3, 1, 1, 1> yyy; swz<T, 3, 2, 2, 2> zzz;
    swz<T, 3, 0, 1, 2> rgb; swz<T, 3, 0, 1, 2> xyz;
  };
  vec() : d{} {}
  template <class A, class B, class C> vec(A a, B b, C c) : d{(T)a, (T)b, (T)c} {}
  template <class S, class = std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, T>>>
  vec(S s) : d{(T)s, (T)s, (T)s} {}
  template <class U> explicit vec(const vec<U, 3>& o) : d{(T)o.d[0], (T)o.d[1], (T)o.d[2]} {}
  template <class S> vec(const vec<T, 2>& a, S c) : d{a.d[0], a.d[1], (T)c} {}
  vec(const vec& o) { for (int i = 0; i < 3; i++) d[i] = o.d[i]; }
  vec& operator=(const vec& o) { for (int i = 0; i < 3; i++) d[i] = o.d[i]; return *this; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
};

template <class T> struct vec<T, 4> {
  union {
    T d[4];
    struct { T x, y, z, w; };
    struct { T r, g, b, a; };
    swz<T, 4, 0, 1> xy; swz<T, 4, 2, 3> zw; swz<T, 4, 0, 1> rg; swz<T, 4, 2, 3> ba;
    swz<T, 4, 0, 2> xz; swz<T, 4, 1, 3> yw;
    swz<T, 4, 0, 0> xx; swz<T, 4, 1, 1> yy; swz<T, 4, 2, 2> zz; swz<T, 4, 3, 3> ww;
    swz<T, 4, 0, 1, 2> rgb; swz<T, 4, 0, 1, 2> xyz;
    swz<T, 4, 0, 0, 0> xxx; swz<T, 4, 1, 1, 1> yyy; swz<T, 4, 2, 2, 2> zzz; swz<T, 4, 3, 3, 3> www;
  };
  vec() : d{} {}
  template <class A, class B, class C, class D> vec(A a, B b, C c, D e) : d{(T)a, (T)b, (T)c, (T)e} {}
  template <class S, class = std::enable_if_t<std::is_arithmetic_v<S> || std::is_same_v<S, T>>>
  vec(S s) : d{(T)s, (T)s, (T)s, (T)s} {}
  template <class U> explicit vec(const vec<U, 4>& o) : d{(T)o.d[0], (T)o.d[1], (T)o.d[2], (T)o.d[3]} {}
  template <class S> vec(const vec<T, 3>& a, S e) : d{a.d[0], a.d[1], a.d[2], (T)e} {}
  vec(const vec<T, 2>& a, const vec<T, 2>& b) : d{a.d[0], a.d[1], b.d[0], b.d[1]} {}
  vec(const vec& o) { for (int i = 0; i < 4; i++) d[i] = o.d[i]; }
  vec& operator=(const vec& o) { for (int i = 0; i < 4; i++) d[i] = o.d[i]; return *this; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
};

// ---- traits: "vec-like" = vec or swizzle proxy ------------------------------------------------
template <class X> struct vl { static constexpr bool is = false; };
template <class T, int N> struct vl<vec<T, N>> {
  static constexpr bool is = true; using E = T; static constexpr int n = N;
  static vec<T, N> get(const vec<T, N>& v) { return v; }
};
template <class T, int N, int... I> struct vl<swz<T, N, I...>> {
  static constexpr bool is = true; using E = T; static constexpr int n = sizeof...(I);
  static vec<T, n> get(const swz<T, N, I...>& v) { return v; }
};
template <class X> constexpr bool is_vl = vl<std::decay_t<X>>::is;
template <class X> constexpr bool is_sc = std::is_arithmetic_v<std::decay_t<X>>
#ifdef FSR1_REF_HALF
                                          || std::is_same_v<std::decay_t<X>, _Float16>
#endif
    ;

// Result element/width of a binary op; scalar operands adopt the vector's element type.
template <class A, class B, bool va = is_vl<A>, bool vb = is_vl<B>> struct bin;
template <class A, class B> struct bin<A, B, true, true> {
  using E = typename vl<std::decay_t<A>>::E; static constexpr int n = vl<std::decay_t<A>>::n;
  static vec<E, n> a(const A& x) { return vl<std::decay_t<A>>::get(x); }
  static vec<E, n> b(const B& x) { return vec<E, n>(vl<std::decay_t<B>>::get(x)); }
};
template <class A, class B> struct bin<A, B, true, false> {
  using E = typename vl<std::decay_t<A>>::E; static constexpr int n = vl<std::decay_t<A>>::n;
  static vec<E, n> a(const A& x) { return vl<std::decay_t<A>>::get(x); }
  static vec<E, n> b(const B& x) { return vec<E, n>((E)x); }
};
template <class A, class B> struct bin<A, B, false, true> {
  using E = typename vl<std::decay_t<B>>::E; static constexpr int n = vl<std::decay_t<B>>::n;
  static vec<E, n> a(const A& x) { return vec<E, n>((E)x); }
  static vec<E, n> b(const B& x) { return vl<std::decay_t<B>>::get(x); }
};
template <class A, class B>
constexpr bool binok = (is_vl<A> && (is_vl<B> || is_sc<B>)) || (is_sc<A> && is_vl<B>);

#define HLSL_BINOP(OP)                                                                   \
  template <class A, class B, class = std::enable_if_t<binok<A, B>>>                     \
  auto operator OP(const A& x, const B& y) {                                             \
    using R = bin<A, B>; auto a = R::a(x); auto b = R::b(y);                             \
    vec<typename R::E, R::n> r;                                                          \
    for (int i = 0; i < R::n; i++) r.d[i] = (typename R::E)(a.d[i] OP b.d[i]);           \
    return r;                                                                            \
  }
HLSL_BINOP(+) HLSL_BINOP(-) HLSL_BINOP(*) HLSL_BINOP(/)
#undef HLSL_BINOP
template <class A, class B, bool ok = binok<A, B>> struct intok { static constexpr bool v = false; };
template <class A, class B> struct intok<A, B, true> {
  static constexpr bool v = std::is_integral_v<typename bin<A, B>::E>;
};
#define HLSL_BITOP(OP)                                                                   \
  template <class A, class B, class = std::enable_if_t<intok<A, B>::v>>                  \
  auto operator OP(const A& x, const B& y) {                                             \
    using R = bin<A, B>; auto a = R::a(x); auto b = R::b(y);                             \
    vec<typename R::E, R::n> r;                                                          \
    for (int i = 0; i < R::n; i++) r.d[i] = (typename R::E)(a.d[i] OP b.d[i]);           \
    return r;                                                                            \
  }
HLSL_BITOP(&) HLSL_BITOP(|) HLSL_BITOP(^) HLSL_BITOP(>>) HLSL_BITOP(<<)
#undef HLSL_BITOP

#define HLSL_CMP(OP)                                                                     \
  template <class A, class B, class = std::enable_if_t<binok<A, B>>>                     \
  auto operator OP(const A& x, const B& y) {                                             \
    using R = bin<A, B>; auto a = R::a(x); auto b = R::b(y);                             \
    vec<bool, R::n> r;                                                                   \
    for (int i = 0; i < R::n; i++) r.d[i] = a.d[i] OP b.d[i];                            \
    return r;                                                                            \
  }
HLSL_CMP(<) HLSL_CMP(>) HLSL_CMP(<=) HLSL_CMP(>=) HLSL_CMP(==) HLSL_CMP(!=)
#undef HLSL_CMP

#define HLSL_ASSIGNOP(OP, BOP)                                                           \
  template <class T, int N, class B, class = std::enable_if_t<is_vl<B> || is_sc<B>>>    \
  vec<T, N>& operator OP(vec<T, N>& x, const B& y) { x = vec<T, N>(x BOP y); return x; } \
  template <class T, int N, int... I, class B, class = std::enable_if_t<is_vl<B> || is_sc<B>>> \
  swz<T, N, I...>& operator OP(swz<T, N, I...>& x, const B& y) {                         \
    x = vec<T, sizeof...(I)>(x BOP y); return x; }
HLSL_ASSIGNOP(+=, +) HLSL_ASSIGNOP(-=, -) HLSL_ASSIGNOP(*=, *) HLSL_ASSIGNOP(/=, /)
HLSL_ASSIGNOP(&=, &) HLSL_ASSIGNOP(|=, |) HLSL_ASSIGNOP(^=, ^) HLSL_ASSIGNOP(>>=, >>) HLSL_ASSIGNOP(<<=, <<)
#undef HLSL_ASSIGNOP

template <class A, class = std::enable_if_t<is_vl<A>>> auto operator-(const A& x) {
  auto a = vl<std::decay_t<A>>::get(x); decltype(a) r;
  for (int i = 0; i < vl<std::decay_t<A>>::n; i++) r.d[i] = (typename vl<std::decay_t<A>>::E)(-a.d[i]);
  return r;
}
template <class A, class = std::enable_if_t<is_vl<A>>> auto operator~(const A& x) {
  auto a = vl<std::decay_t<A>>::get(x); decltype(a) r;
  for (int i = 0; i < vl<std::decay_t<A>>::n; i++) r.d[i] = (typename vl<std::decay_t<A>>::E)(~a.d[i]);
  return r;
}

// ---- scalar intrinsics ------------------------------------------------------------------------
inline float s_min(float a, float b) { return fminf(a, b); }
inline float s_max(float a, float b) { return fmaxf(a, b); }
inline uint s_min(uint a, uint b) { return a < b ? a : b; }
inline uint s_max(uint a, uint b) { return a > b ? a : b; }
inline int s_min(int a, int b) { return a < b ? a : b; }
inline int s_max(int a, int b) { return a > b ? a : b; }
inline float s_abs(float a) { return fabsf(a); }
inline int s_abs(int a) { return a < 0 ? -a : a; }
inline float s_floor(float a) { return floorf(a); }
inline float s_sqrt(float a) { return sqrtf(a); }
inline float s_rcp(float a) { return 1.0f / a; }
inline float s_rsqrt(float a) { return 1.0f / sqrtf(a); }
inline float s_sat(float a) { return fminf(fmaxf(a, 0.0f), 1.0f); }
inline float s_sin(float a) { return sinf(a); }
inline float s_cos(float a) { return cosf(a); }
inline float s_exp2(float a) { return exp2f(a); }
inline float s_log2(float a) { return log2f(a); }
inline float s_trunc(float a) { return truncf(a); }
inline float s_pow(float a, float b) { return powf(a, b); }
#ifdef FSR1_REF_HALF
// IEEE minNum/maxNum on halves, exact because every half is a float.
inline _Float16 s_min(_Float16 a, _Float16 b) { return (_Float16)fminf((float)a, (float)b); }
inline _Float16 s_max(_Float16 a, _Float16 b) { return (_Float16)fmaxf((float)a, (float)b); }
inline _Float16 s_abs(_Float16 a) { return (_Float16)fabsf((float)a); }
inline _Float16 s_floor(_Float16 a) { return (_Float16)floorf((float)a); }
inline _Float16 s_sqrt(_Float16 a) { return (_Float16)sqrtf((float)a); }
inline _Float16 s_rcp(_Float16 a) { return (_Float16)1.0 / a; }
inline _Float16 s_rsqrt(_Float16 a) { return (_Float16)(1.0f / sqrtf((float)a)); }
inline _Float16 s_sat(_Float16 a) { return (_Float16)fminf(fmaxf((float)a, 0.0f), 1.0f); }
inline _Float16 s_sin(_Float16 a) { return (_Float16)sinf((float)a); }
inline _Float16 s_cos(_Float16 a) { return (_Float16)cosf((float)a); }
inline _Float16 s_exp2(_Float16 a) { return (_Float16)exp2f((float)a); }
inline _Float16 s_log2(_Float16 a) { return (_Float16)log2f((float)a); }
inline _Float16 s_trunc(_Float16 a) { return (_Float16)truncf((float)a); }
inline _Float16 s_pow(_Float16 a, _Float16 b) { return (_Float16)powf((float)a, (float)b); }
inline uint16_t s_min(uint16_t a, uint16_t b) { return a < b ? a : b; }
inline uint16_t s_max(uint16_t a, uint16_t b) { return a > b ? a : b; }
inline int16_t s_min(int16_t a, int16_t b) { return a < b ? a : b; }
inline int16_t s_max(int16_t a, int16_t b) { return a > b ? a : b; }
inline int16_t s_abs(int16_t a) { return a < 0 ? (int16_t)-a : a; }
#endif

inline uint s_asuint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline uint s_asuint(uint u) { return u; }
inline uint s_asuint(int i) { return (uint)i; }
inline float s_asfloat(uint u) { float f; memcpy(&f, &u, 4); return f; }
inline float s_asfloat(float f) { return f; }
inline float s_asfloat(int i) { return s_asfloat((uint)i); }

// IEEE binary32 -> binary16, round to nearest even (HLSL f32tof16); returns the 16 bits.
inline uint s_f32tof16(float f) {
  uint x = s_asuint(f), sign = (x >> 16) & 0x8000u; x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u);
  if (x >= 0x477ff000u) return sign | 0x7c00u;                     // rounds to >= 65520 -> inf
  if (x < 0x33000001u) return sign;                                // < 2^-25 (or tie) -> 0
  int e = (int)(x >> 23) - 127; uint m = (x & 0x7fffffu) | 0x800000u;
  int shift = e < -14 ? (13 + (-14 - e)) : 13;                      // subnormal halves shift more
  uint half = e < -14 ? 0u : (uint)(e + 15) << 10;
  uint q = m >> shift, rem = m & ((1u << shift) - 1u), mid = 1u << (shift - 1);
  if (e >= -14) q &= 0x3ffu;
  uint h = half + q;
  if (rem > mid || (rem == mid && (h & 1u))) h++;
  return sign | h;
}
inline float s_f16tof32(uint h) {
  uint sign = (h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 0x3ffu;
  if (e == 31) return s_asfloat(sign | 0x7f800000u | (m << 13));
  if (e == 0) { float v = (float)m * 5.9604644775390625e-8f; return sign ? -v : v; }
  return s_asfloat(sign | ((e + 112u) << 23) | (m << 13));
}

}  // namespace hlsl

// ---- the HLSL-visible names ---------------------------------------------------------------------
#define HLSL_TYPES(T)                                                         \
  typedef hlsl::vec<T, 2> T##2; typedef hlsl::vec<T, 3> T##3; typedef hlsl::vec<T, 4> T##4;
HLSL_TYPES(float) HLSL_TYPES(uint) HLSL_TYPES(int) HLSL_TYPES(bool)
#ifdef FSR1_REF_HALF
HLSL_TYPES(min16float) HLSL_TYPES(min16uint) HLSL_TYPES(min16int)
#endif
#undef HLSL_TYPES

#define HLSL_FN1(NAME, SFN)                                                                   \
  template <class A, class = std::enable_if_t<hlsl::is_vl<A>>> auto NAME(const A& x) {          \
    auto a = hlsl::vl<std::decay_t<A>>::get(x); decltype(a) r;                                  \
    for (int i = 0; i < hlsl::vl<std::decay_t<A>>::n; i++) r.d[i] = hlsl::SFN(a.d[i]);          \
    return r; }
#define HLSL_FN2(NAME, SFN)                                                                   \
  template <class A, class B, class = std::enable_if_t<hlsl::binok<A, B>>>                     \
  auto NAME(const A& x, const B& y) {                                                          \
    using R = hlsl::bin<A, B>; auto a = R::a(x); auto b = R::b(y);                              \
    hlsl::vec<typename R::E, R::n> r;                                                          \
    for (int i = 0; i < R::n; i++) r.d[i] = hlsl::SFN(a.d[i], b.d[i]);                         \
    return r; }

#define HLSL_SCALAR1(NAME, SFN, T) inline T NAME(T a) { return hlsl::SFN(a); }
#define HLSL_SCALAR2(NAME, SFN, T) inline T NAME(T a, T b) { return hlsl::SFN(a, b); }

#ifdef FSR1_REF_HALF
#define HLSL_ALLF1(NAME, SFN) HLSL_FN1(NAME, SFN) HLSL_SCALAR1(NAME, SFN, float) HLSL_SCALAR1(NAME, SFN, _Float16)
#define HLSL_ALLF2(NAME, SFN) HLSL_FN2(NAME, SFN) HLSL_SCALAR2(NAME, SFN, float) HLSL_SCALAR2(NAME, SFN, _Float16)
#else
#define HLSL_ALLF1(NAME, SFN) HLSL_FN1(NAME, SFN) HLSL_SCALAR1(NAME, SFN, float)
#define HLSL_ALLF2(NAME, SFN) HLSL_FN2(NAME, SFN) HLSL_SCALAR2(NAME, SFN, float)
#endif

using std::abs;  // int abs from <cstdlib>/<cmath>; float/half/vec versions below
HLSL_FN1(abs, s_abs)
#ifdef FSR1_REF_HALF
inline _Float16 abs(_Float16 a) { return hlsl::s_abs(a); }
inline int16_t abs(int16_t a) { return hlsl::s_abs(a); }
#endif
HLSL_ALLF1(floor, s_floor) HLSL_ALLF1(sqrt, s_sqrt) HLSL_ALLF1(rcp, s_rcp) HLSL_ALLF1(rsqrt, s_rsqrt)
HLSL_ALLF1(saturate, s_sat) HLSL_ALLF1(sin, s_sin) HLSL_ALLF1(cos, s_cos) HLSL_ALLF1(exp2, s_exp2)
HLSL_ALLF1(log2, s_log2) HLSL_ALLF1(trunc, s_trunc)
HLSL_ALLF2(pow, s_pow)
HLSL_ALLF2(min, s_min) HLSL_ALLF2(max, s_max)
HLSL_SCALAR2(min, s_min, uint) HLSL_SCALAR2(max, s_max, uint)
HLSL_SCALAR2(min, s_min, int) HLSL_SCALAR2(max, s_max, int)
#ifdef FSR1_REF_HALF
HLSL_SCALAR2(min, s_min, uint16_t) HLSL_SCALAR2(max, s_max, uint16_t)
HLSL_SCALAR2(min, s_min, int16_t) HLSL_SCALAR2(max, s_max, int16_t)
#endif
#define HLSL_FN1T(NAME, SFN, RT)                                                             \
  template <class A, class = std::enable_if_t<hlsl::is_vl<A>>> auto NAME(const A& x) {          \
    auto a = hlsl::vl<std::decay_t<A>>::get(x); hlsl::vec<RT, hlsl::vl<std::decay_t<A>>::n> r;  \
    for (int i = 0; i < hlsl::vl<std::decay_t<A>>::n; i++) r.d[i] = hlsl::SFN(a.d[i]);          \
    return r; }
HLSL_FN1T(asuint, s_asuint, uint) HLSL_FN1T(asfloat, s_asfloat, float)
HLSL_FN1T(f32tof16, s_f32tof16, uint) HLSL_FN1T(f16tof32, s_f16tof32, float)
inline uint asuint(float a) { return hlsl::s_asuint(a); }
inline uint asuint(uint a) { return a; }
inline uint asuint(int a) { return (uint)a; }
inline float asfloat(uint a) { return hlsl::s_asfloat(a); }
inline float asfloat(int a) { return hlsl::s_asfloat((uint)a); }
inline float asfloat(float a) { return a; }
inline uint f32tof16(float a) { return hlsl::s_f32tof16(a); }
inline float f16tof32(uint a) { return hlsl::s_f16tof32(a); }

// lerp(x,y,s) = x + s*(y-x) ; clamp(x,lo,hi) = min(max(x,lo),hi)
template <class A, class B, class C> auto lerp(const A& x, const B& y, const C& s) { return x + s * (y - x); }
template <class A, class B, class C> auto clamp(const A& x, const B& lo, const C& hi) { return min(max(x, lo), hi); }
template <class A, class B, class = std::enable_if_t<hlsl::is_vl<A> && hlsl::is_vl<B>>>
auto dot(const A& x, const B& y) {
  using R = hlsl::bin<A, B>; auto a = R::a(x); auto b = R::b(y);
  typename R::E s = a.d[0] * b.d[0];
  for (int i = 1; i < R::n; i++) s = s + a.d[i] * b.d[i];
  return s;
}
