/**
 * \file host_kernels.cc
 * \brief CPU twins of the data-plane kernels (see host_kernels.h).
 *
 * Same wire formats and the same rounding decisions as src/kernels/copy_kernels.cu /
 * update_kernels.cu, written as plain loops split over a few threads:
 *   bf16        round-to-nearest-even of the fp32 value (NaN stays NaN)
 *   fp8-block   32 elements share one e8m0 exponent chosen so that amax maps into
 *               [224, 448] of e4m3; values are rounded to nearest-even and saturate at 448
 * The optimizer step uses IEEE sqrt / division where the GPU kernel uses MUFU
 * approximations, so the two agree to a few ulp, not bit for bit.
 */
#include "kernels/host_kernels.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint32_t F2U(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
inline float U2F(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline uint16_t F32ToBf16(float f) {
  uint32_t u = F2U(f);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // quiet NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
inline float Bf16ToF32(uint16_t h) { return U2F(static_cast<uint32_t>(h) << 16); }

/*! \brief e4m3 (bias 7, no infinities, 0x7f = NaN) -> fp32, by table */
const float* E4m3Table() {
  static float table[256];
  static bool ready = [] {
    for (int b = 0; b < 256; ++b) {
      const int sign = b >> 7, exp = (b >> 3) & 0xf, man = b & 7;
      float v;
      if (exp == 0) {
        v = std::ldexp(static_cast<float>(man), -9);  // subnormal: man/8 * 2^-6
      } else if (exp == 15 && man == 7) {
        v = std::nanf("");
      } else {
        v = std::ldexp(1.f + man / 8.f, exp - 7);
      }
      table[b] = sign ? -v : v;
    }
    return true;
  }();
  (void)ready;
  return table;
}

/*! \brief fp32 -> e4m3, round to nearest even, saturating at +-448 (cvt.rn.satfinite) */
inline uint8_t F32ToE4m3(float f) {
  const uint32_t u = F2U(f);
  const uint8_t sign = static_cast<uint8_t>((u >> 24) & 0x80);
  const float a = std::fabs(f);
  if (a != a) return static_cast<uint8_t>(sign | 0x7f);
  if (a >= 448.f) return static_cast<uint8_t>(sign | 0x7e);
  if (a < std::ldexp(1.f, -10)) return sign;  // below half of the smallest subnormal (2^-9)
  int e;
  std::frexp(a, &e);  // a = m * 2^e, m in [0.5, 1)
  int exp = e - 1;    // a = 1.xxx * 2^exp
  if (exp < -6) exp = -6;  // subnormal range shares the exponent of 2^-6
  // quantum of the 3-bit mantissa at this exponent
  const float q = std::ldexp(1.f, exp - 3);
  float r = std::nearbyint(a / q);  // current rounding mode is nearest-even
  float v = r * q;
  if (v >= 448.f) return static_cast<uint8_t>(sign | 0x7e);
  if (v < std::ldexp(1.f, -6)) {  // subnormal result: mantissa = v / 2^-9
    return static_cast<uint8_t>(sign | static_cast<uint8_t>(std::lrint(std::ldexp(v, 9))));
  }
  std::frexp(v, &e);
  exp = e - 1;
  const int man = static_cast<int>(std::lrint((std::ldexp(v, -exp) - 1.f) * 8.f));
  return static_cast<uint8_t>(sign | ((exp + 7) << 3) | man);
}

inline uint32_t E8m0ForAmax(float amax) {
  if (!(amax > 0.f)) return 0u;
  const uint32_t bits = F2U(amax);
  const int k = static_cast<int>((bits >> 23) & 0xff) - 127;
  const uint32_t mant = bits & 0x7fffffu;
  int e = k - 8 + (mant > 0x600000u ? 1 : 0);  // 448 = 1.75 * 2^8
  e = std::max(-127, std::min(127, e));
  return static_cast<uint32_t>(e + 127);
}
inline float Exp2FromE8m0(uint32_t biased) {
  const int be = std::max(1, std::min(254, static_cast<int>(biased)));
  return U2F(static_cast<uint32_t>(be) << 23);
}
inline float Exp2FromE8m0Neg(uint32_t biased) {
  const int e = 127 - static_cast<int>(biased);
  const int be = std::max(1, std::min(254, e + 127));
  return U2F(static_cast<uint32_t>(be) << 23);
}

/*! \brief run fn(begin, end) over [0, n) in `grain`-aligned pieces on a few threads */
template <typename Fn>
void ParallelFor(size_t n, size_t grain, Fn fn) {
  const size_t kMinPerThread = 1u << 18;
  size_t threads = std::min<size_t>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
  threads = std::min(threads, std::max<size_t>(1, n / kMinPerThread));
  if (threads <= 1) {
    fn(static_cast<size_t>(0), n);
    return;
  }
  size_t chunk = (n + threads - 1) / threads;
  chunk = (chunk + grain - 1) / grain * grain;
  std::vector<std::thread> pool;
  for (size_t a = chunk; a < n; a += chunk) {
    const size_t b = std::min(n, a + chunk);
    pool.emplace_back([=] { fn(a, b); });
  }
  fn(static_cast<size_t>(0), std::min(n, chunk));
  for (auto& t : pool) t.join();
}

template <bool SRC_BF16>
void QuantFp8Block(unsigned char* payload, unsigned char* scales, const void* src, size_t n,
                   float scale) {
  const size_t blocks = (n + 31) / 32;
  ParallelFor(blocks, 1, [=](size_t b0, size_t b1) {
    for (size_t b = b0; b < b1; ++b) {
      float x[32];
      float amax = 0.f;
      for (int j = 0; j < 32; ++j) {
        const size_t e = b * 32 + j;
        float v = 0.f;
        if (e < n) {
          v = SRC_BF16 ? Bf16ToF32(static_cast<const uint16_t*>(src)[e])
                       : static_cast<const float*>(src)[e];
        }
        x[j] = v * scale;
        amax = std::max(amax, std::fabs(x[j]));
      }
      const uint32_t eb = E8m0ForAmax(amax);
      const float inv = Exp2FromE8m0Neg(eb);
      for (int j = 0; j < 32; ++j) payload[b * 32 + j] = F32ToE4m3(x[j] * inv);
      scales[b] = static_cast<unsigned char>(eb);
    }
  });
}

/*! \brief value of element e of a gradient slot in format fmt */
inline float LoadGrad(const void* slot, int fmt, size_t e, size_t npad, const float* e4m3) {
  const unsigned char* base = static_cast<const unsigned char*>(slot);
  switch (fmt) {
    case PS_GRAD_BF16:
      return Bf16ToF32(reinterpret_cast<const uint16_t*>(base)[e]);
    case PS_GRAD_FP8BLOCK:
      return e4m3[base[e]] * Exp2FromE8m0(base[npad + (e >> 5)]);
    default:
      return reinterpret_cast<const float*>(base)[e];
  }
}

}  // namespace

extern "C" int ps_host_copy(void* dst, const void* src, size_t n, int codec, float scale) {
  if (n == 0) return 0;
  switch (codec) {
    case PS_CODEC_RAW:
      if (dst != src) memmove(dst, src, n);
      return 0;
    case PS_CODEC_F32_TO_BF16: {
      const float* s = static_cast<const float*>(src);
      uint16_t* d = static_cast<uint16_t*>(dst);
      ParallelFor(n / 4, 1, [=](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) d[i] = F32ToBf16(s[i] * scale);
      });
      return 0;
    }
    case PS_CODEC_BF16_SCALE: {
      const uint16_t* s = static_cast<const uint16_t*>(src);
      uint16_t* d = static_cast<uint16_t*>(dst);
      ParallelFor(n / 2, 1, [=](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) d[i] = F32ToBf16(Bf16ToF32(s[i]) * scale);
      });
      return 0;
    }
    case PS_CODEC_F32_TO_FP8BLOCK:
    case PS_CODEC_BF16_TO_FP8BLOCK: {
      const bool bf = codec == PS_CODEC_BF16_TO_FP8BLOCK;
      const size_t ne = n / (bf ? 2 : 4);
      const size_t npad = (ne + 31) / 32 * 32;
      unsigned char* payload = static_cast<unsigned char*>(dst);
      if (bf) QuantFp8Block<true>(payload, payload + npad, src, ne, scale);
      else QuantFp8Block<false>(payload, payload + npad, src, ne, scale);
      return 0;
    }
    default:
      return 1;
  }
}

extern "C" int ps_host_decode(float* dst, const void* wire, size_t n, int fmt) {
  if (n == 0) return 0;
  if (fmt != PS_GRAD_BF16 && fmt != PS_GRAD_FP8BLOCK && fmt != PS_GRAD_F32) return 1;
  const size_t npad = (n + 31) / 32 * 32;
  const float* e4m3 = E4m3Table();
  ParallelFor(n, 32, [=](size_t a, size_t b) {
    for (size_t i = a; i < b; ++i) dst[i] = LoadGrad(wire, fmt, i, npad, e4m3);
  });
  return 0;
}

extern "C" int ps_host_sum(float* out, const void* const* grads, int num_grads, int fmt, size_t n,
                           float scale, int accumulate) {
  if (n == 0) return 0;
  if (num_grads < 1 || num_grads > PS_MAX_FANIN) return 1;
  if (fmt != PS_GRAD_BF16 && fmt != PS_GRAD_FP8BLOCK && fmt != PS_GRAD_F32) return 1;
  const size_t npad = (n + 31) / 32 * 32;
  const float* e4m3 = E4m3Table();
  std::vector<const void*> g(grads, grads + num_grads);
  ParallelFor(n, 32, [=, &g](size_t a, size_t b) {
    for (size_t i = a; i < b; ++i) {
      float s = 0.f;
      for (const void* slot : g) s += LoadGrad(slot, fmt, i, npad, e4m3);
      out[i] = (accumulate ? out[i] : 0.f) + s * scale;
    }
  });
  return 0;
}

extern "C" int ps_host_update(const ps_update_args* a, const ps_opt_params* o) {
  if (a->n == 0) return 0;
  if (a->num_grads < 1 || a->num_grads > PS_MAX_FANIN) return 1;
  if (a->num_outs < 0 || a->num_outs > PS_MAX_FANOUT) return 1;
  const int fmt = a->grad_format;
  if (fmt != PS_GRAD_BF16 && fmt != PS_GRAD_FP8BLOCK && fmt != PS_GRAD_F32) return 1;
  const size_t n = a->n, npad = (n + 31) / 32 * 32;
  const float* e4m3 = E4m3Table();
  const ps_update_args args = *a;
  const ps_opt_params opt = *o;
  const bool adam = opt.optimizer == PS_OPT_ADAMW;
  ParallelFor(n, 32, [=](size_t b0, size_t b1) {
    for (size_t i = b0; i < b1; ++i) {
      float g = 0.f;
      for (int w = 0; w < args.num_grads; ++w) g += LoadGrad(args.grads[w], fmt, i, npad, e4m3);
      g *= opt.grad_scale;
      float p = args.master[i], m = args.m[i];
      if (adam) {
        float v = args.v[i];
        m = opt.beta1 * m + (1.f - opt.beta1) * g;
        v = opt.beta2 * v + (1.f - opt.beta2) * g * g;
        const float denom = std::sqrt(v / opt.bias_corr2) + opt.eps;
        p = p - opt.lr * ((m / opt.bias_corr1) / denom + opt.weight_decay * p);
        args.v[i] = v;
      } else {
        g += opt.weight_decay * p;
        m = opt.beta1 * m + g;
        p = p - opt.lr * m;
      }
      args.master[i] = p;
      args.m[i] = m;
      for (int k = 0; k < args.num_outs; ++k) {
        if (args.out_f32) static_cast<float*>(args.outs[k])[i] = p;
        else static_cast<uint16_t*>(args.outs[k])[i] = F32ToBf16(p);
      }
    }
  });
  return 0;
}
