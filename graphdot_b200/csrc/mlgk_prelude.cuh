// mlgk_prelude.cuh -- fixed device-side vocabulary available to the spliced
// microkernel expressions.  Compiled by NVRTC (no system headers) and by nvcc.
//
// Provides what the reference's generated expressions may name (SURVEY.md
// appendix A; reference graphdot/cpp/numpy_type.h, fmath.h, frozen_array.h,
// basekernel/{normalize,convolution,dotproduct}.h): numpy scalar aliases,
// graphdot::ipow<N>/ripow<N>, frozen_array<T>, normalize / normalize_jacobian,
// convolution<mean> / convolution_jacobian<mean>, dotproduct.
#pragma once

typedef bool bool_;
typedef signed char int8;
typedef short int16;
typedef int int32;
typedef long long int64;
typedef unsigned char uint8;
typedef unsigned short uint16;
typedef unsigned int uint32;
typedef unsigned long long uint64;
typedef float float32;
typedef double float64;
typedef long long intp;
typedef unsigned long long uintp;
struct empty_t {};

namespace graphdot {

// x^E for a compile-time non-negative integer E by repeated squaring.
template<int E, class F> __host__ __device__ __forceinline__ constexpr F ipow(F base) {
    if constexpr (E == 0) {
        return F(1);
    } else if constexpr (E == 1) {
        return base;
    } else {
        F h = ipow<E / 2>(base);
        return (E & 1) ? h * h * base : h * h;
    }
}

// x^-E
template<int E, class F> __host__ __device__ __forceinline__ constexpr F ripow(F base) {
    return ipow<E>(F(1) / base);
}

}  // namespace graphdot

// Read-only view of a variable-length feature; `_data` holds an absolute
// device address after upload (pool-relative offset inside a packed blob).
template<class T> struct frozen_array {
    const T *_data;
    int32 size;
    __device__ __forceinline__ const T *begin() const { return _data; }
    __device__ __forceinline__ const T *end() const { return _data + size; }
    __device__ __forceinline__ const T &operator[](int i) const { return _data[i]; }
};

// k(x,y) / sqrt(k(x,x) k(y,y)); 0 when a self-similarity is not positive.
template<class F, class X, class Y>
__device__ __forceinline__ float normalize(F const f, X const &x, Y const &y) {
    float const kxx = f(x, x), kyy = f(y, y);
    float const s = kxx * kyy;
    return s > 0.f ? f(x, y) * rsqrtf(s) : 0.f;
}

template<class F, class J, class X, class Y>
__device__ __forceinline__ float normalize_jacobian(F const f, J const j, X const &x, Y const &y) {
    float const kxx = f(x, x), kxy = f(x, y), kyy = f(y, y);
    float const jxx = j(x, x), jxy = j(x, y), jyy = j(y, y);
    float const s = kxx * kyy;
    if (!(s > 0.f)) return 0.f;
    float const rs = rsqrtf(s);
    return jxy * rs - 0.5f * kxy * rs * rs * rs * (jxx * kyy + kxx * jyy);
}

// sum (or mean) of f over all element pairs of two sequences.  The elements of y
// are held in registers four at a time while x streams by: 4 + |x| loads per chunk
// instead of one load per evaluation (an 8 x 8 convolution is evaluated once per
// product-graph node; its loads are scattered over the feature pools of 32 nodes).
template<bool mean, class F, class X, class Y>
__device__ __forceinline__ float convolution(F const f, X const &x, Y const &y) {
    float k = 0.f;
    const int nx = x.size, ny = y.size;
    for (int j0 = 0; j0 < ny; j0 += 4) {
        const int m = ny - j0;  // elements in this chunk (at least 1)
        const auto y0 = y[j0], y1 = y[j0 + (m > 1 ? 1 : 0)], y2 = y[j0 + (m > 2 ? 2 : 0)], y3 = y[j0 + (m > 3 ? 3 : 0)];
        for (int a = 0; a < nx; ++a) {
            const auto xa = x[a];
            float t = f(xa, y0);
            if (m > 1) t += f(xa, y1);
            if (m > 2) t += f(xa, y2);
            if (m > 3) t += f(xa, y3);
            k += t;
        }
    }
    return mean ? k / (float)(nx * ny) : k;
}

template<bool mean, class J, class X, class Y>
__device__ __forceinline__ float convolution_jacobian(J const j, X const &x, Y const &y) {
    return convolution<mean>(j, x, y);
}

template<class T>
__device__ __forceinline__ float dotproduct(frozen_array<T> const &x, frozen_array<T> const &y) {
    float s = 0.f;
    for (int i = 0; i < x.size; ++i) s += (float)x._data[i] * (float)y._data[i];
    return s;
}
