// oracle/stubs/trng/*.hpp -- TEST INFRASTRUCTURE ONLY: stand-in for the TRNG engines/distributions named by
// SimToolbox/Util/TRngPool.hpp (TRNG is absent from this image).  NOT TRNG's streams: a splitmix64 counter generator
// with leap-frog split(); u01 = 53 random bits, n01 = Box-Muller (cosine branch, two uniforms per deviate).  Tests that
// need the deviates the reference drew read them back through oracle/ref_system_driver.cpp.
#pragma once
#include <cmath>
#include <cstdint>
namespace trng {
class mrg5 {
  public:
    mrg5() : ctr_(0), stride_(1) {}
    void seed(unsigned long s) { ctr_ = (uint64_t)s * 0x9E3779B97F4A7C15ull; }
    void split(unsigned int s, unsigned int n) { // leap-frog: stream n of s
        ctr_ += (uint64_t)n * stride_;
        stride_ *= s;
    }
    void jump(unsigned long long k) { ctr_ += k * stride_; }
    uint64_t operator()() {
        uint64_t z = (ctr_ + 0x9E3779B97F4A7C15ull);
        ctr_ += stride_;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }

  private:
    uint64_t ctr_, stride_;
};
typedef mrg5 lcg64_shift_stub_base;
template <class T>
class uniform01_dist {
  public:
    template <class E>
    T operator()(E &e) { return (T)((e() >> 11) * (1.0 / 9007199254740992.0)); }
};
template <class T>
class normal_dist {
  public:
    normal_dist(T mu, T sigma) : mu_(mu), sigma_(sigma) {}
    template <class E>
    T operator()(E &e) {
        uniform01_dist<T> u;
        T a = u(e), b = u(e);
        if (a < 1e-300) a = 1e-300;
        return mu_ + sigma_ * std::sqrt(-2.0 * std::log(a)) * std::cos(6.283185307179586476925286766559 * b);
    }

  private:
    T mu_, sigma_;
};
template <class T>
class lognormal_dist {
  public:
    lognormal_dist(T mu, T sigma) : n_(mu, sigma) {}
    template <class E>
    T operator()(E &e) { return std::exp(n_(e)); }

  private:
    normal_dist<T> n_;
};
} // namespace trng
