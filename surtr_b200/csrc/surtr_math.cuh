// surtr_math.cuh -- the float32 arithmetic contract of the cutting path, as __device__ functions.
//
// The reference computes everything through DirectX::SimpleMath -> DirectXMath (SSE2 code path, no FMA):
// every product and every sum is rounded separately.  Piece-to-cell assignments and
// vertex/face counts only reproduce bit-exactly if the kernels round identically, so every operation here is
// spelled with the never-contracted intrinsics (__fmul_rn / __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn)
// and the translation unit is additionally built with -fmad=false.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace surtr
{
// Vector3::Dot (ThirdParty/Inc/SimpleMath.inl:918-925): (ax*bx + ay*by) + az*bz
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

// Signed distance of Poly::ComparePlanePoint (Src/Poly.cpp:718): plane.D() + plane.Normal().Dot(point)
__device__ __forceinline__ float signed_dist(const float4& pl, float x, float y, float z)
{
    return __fadd_rn(pl.w, dot3(pl.x, pl.y, pl.z, x, y, z));
}

// Poly::ComparePlanePoint (Src/Poly.cpp:716-723): +1 keep, 0 in-plane (|s| < 1e-10, compared as double), -1 clipped.
// A NaN distance (degenerate zero-area cell face -> NaN plane) gives 0, exactly as sgn0(-NaN) does (Poly.cpp:32).
__device__ __forceinline__ int classify(float s)
{
    // (double)|s| < 1e-10 holds exactly for the floats below 0x2EDBE6FF (= 1.0000000134e-10f, the float next above the
    // double 1e-10; 0x2EDBE6FE is the largest float under it) -- one float compare, no FP64 round trip
    if (fabsf(s) < __uint_as_float(0x2EDBE6FFu))
        return 0;
    return s < 0.f ? 1 : (s > 0.f ? -1 : 0);
}

// Poly::PlaneLineIntersection (Src/Poly.cpp:746-751): ((a*sb) - (b*sa)) / (sb - sa), Vector3/float = * (1.f/s)
__device__ __forceinline__ void plane_line_intersection(float ax, float ay, float az, float sa, float bx, float by,
                                                        float bz, float sb, float& ox, float& oy, float& oz)
{
    const float r = __fdiv_rn(1.f, __fsub_rn(sb, sa));
    ox = __fmul_rn(__fsub_rn(__fmul_rn(ax, sb), __fmul_rn(bx, sa)), r);
    oy = __fmul_rn(__fsub_rn(__fmul_rn(ay, sb), __fmul_rn(by, sa)), r);
    oz = __fmul_rn(__fsub_rn(__fmul_rn(az, sb), __fmul_rn(bz, sa)), r);
}

// Vector3::Cross (SimpleMath.inl:936-946)
__device__ __forceinline__ void cross3(float ax, float ay, float az, float bx, float by, float bz, float& ox,
                                       float& oy, float& oz)
{
    ox = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
    oy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
    oz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
}

// Plane(point, normal) (SimpleMath.inl:2782-2788): (n, -Dot(point, n))
__device__ __forceinline__ float4 plane_from_point_normal(float px, float py, float pz, float nx, float ny, float nz)
{
    return make_float4(nx, ny, nz, -dot3(px, py, pz, nx, ny, nz));
}

// Plane(p1, p2, p3) (SimpleMath.inl:2773-2780 = XMPlaneFromPoints): n = normalize((p1 - p2) x (p1 - p3)), d = -n.p1;
// XMVector3Normalize divides by sqrt(dot) and maps zero length to 0.
__device__ __forceinline__ float4 plane_from_points(float ax, float ay, float az, float bx, float by, float bz, float cx, float cy,
                                                    float cz)
{
    float nx, ny, nz;
    cross3(__fsub_rn(ax, bx), __fsub_rn(ay, by), __fsub_rn(az, bz), __fsub_rn(ax, cx), __fsub_rn(ay, cy), __fsub_rn(az, cz), nx, ny, nz);
    const float lsq = dot3(nx, ny, nz, nx, ny, nz);
    if (lsq == 0.f) { nx = ny = nz = 0.f; }
    else
    {
        const float len = __fsqrt_rn(lsq);
        nx = __fdiv_rn(nx, len); ny = __fdiv_rn(ny, len); nz = __fdiv_rn(nz, len);
    }
    return make_float4(nx, ny, nz, -dot3(nx, ny, nz, ax, ay, az));
}
} // namespace surtr
