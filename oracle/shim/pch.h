// oracle/shim/pch.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Headless stand-in for the reference's precompiled header (Inc/pch.h) so that the
// reference's own geometry sources (Src/Poly.cpp, Src/Kdop.cpp, Src/VMACH.cpp,
// Inc/DT3D.h) compile with g++ from where they lie under /root/reference.
// It supplies:
//   * the few Windows typedefs/macros those files use (Inc/pch.h:18 EPSILON,
//     pch.h:103-108 UniqueVector, pch.h:120 OutputDebugStringWFormat),
//   * a restatement of the DirectXMath/SimpleMath value types they touch
//     (ThirdParty/Inc/SimpleMath.inl:729-1006 Vector3, :2773-2788 Plane), with the
//     float32 operation order of the SSE2 code path:
//     every product and sum is rounded separately -- build with -ffp-contract=off,
//   * `#define MESH_H` + a plain VertexNormalColor so Inc/Mesh.h's D3D12 body is skipped.
//
// DirectXMath itself is a Windows-SDK header that is neither vendored in the reference
// nor present in this image; the arithmetic below is therefore the *definition* the
// oracle, the host library and the CUDA kernels all share (parity "unpinned" against
// the MSVC /fp:fast binary).
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <functional>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define EPSILON 1e-12

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif
typedef unsigned int UINT;
typedef int BOOL;
#define _In_
#define _Out_
#define _Inout_

#define OutputDebugStringWFormat(...) ((void)0)

template <typename T>
static void UniqueVector(const std::vector<T>& dupVec, std::vector<T>& uniqueVec)
{
	for (const T& e : dupVec)
		if (uniqueVec.end() == std::find(uniqueVec.begin(), uniqueVec.end(), e))
			uniqueVec.push_back(e);
}

namespace DirectX
{
constexpr float XM_2PI = 6.283185307f;

struct XMFLOAT3
{
	float x, y, z;
	XMFLOAT3() = default;
	constexpr XMFLOAT3(float ix, float iy, float iz) : x(ix), y(iy), z(iz) {}
};

struct XMFLOAT4
{
	float x, y, z, w;
	XMFLOAT4() = default;
	constexpr XMFLOAT4(float ix, float iy, float iz, float iw) : x(ix), y(iy), z(iz), w(iw) {}
};

// Row-major 4x4, rows r[0..3] as in DirectXMath.
struct XMMATRIX
{
	float r[4][4];
};

inline XMMATRIX XMMatrixTranspose(const XMMATRIX& m)
{
	XMMATRIX t;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			t.r[i][j] = m.r[j][i];
	return t;
}

namespace SimpleMath
{
struct Vector3 : public XMFLOAT3
{
	Vector3() : XMFLOAT3(0.f, 0.f, 0.f) {}
	constexpr explicit Vector3(float ix) : XMFLOAT3(ix, ix, ix) {}
	constexpr Vector3(float ix, float iy, float iz) : XMFLOAT3(ix, iy, iz) {}
	// double/int literals narrow exactly like the MSVC build's implicit conversions.
	Vector3(double ix, double iy, double iz) : XMFLOAT3((float)ix, (float)iy, (float)iz) {}
	Vector3(int ix, int iy, int iz) : XMFLOAT3((float)ix, (float)iy, (float)iz) {}
	Vector3(const XMFLOAT3& v) : XMFLOAT3(v.x, v.y, v.z) {}

	bool operator==(const Vector3& v) const { return x == v.x && y == v.y && z == v.z; }
	bool operator!=(const Vector3& v) const { return !(*this == v); }

	Vector3& operator+=(const Vector3& v) { x = x + v.x; y = y + v.y; z = z + v.z; return *this; }
	Vector3& operator-=(const Vector3& v) { x = x - v.x; y = y - v.y; z = z - v.z; return *this; }
	Vector3& operator*=(const Vector3& v) { x = x * v.x; y = y * v.y; z = z * v.z; return *this; }
	Vector3& operator*=(float s) { x = x * s; y = y * s; z = z * s; return *this; }
	// SimpleMath.inl:788-796: XMVectorScale(v, 1.f / S)
	Vector3& operator/=(float s) { const float r = 1.f / s; x = x * r; y = y * r; z = z * r; return *this; }

	Vector3 operator-() const { return Vector3(-x, -y, -z); }

	float Dot(const Vector3& v) const { return (x * v.x + y * v.y) + z * v.z; }
	float LengthSquared() const { return Dot(*this); }
	float Length() const { return std::sqrt(Dot(*this)); }

	Vector3 Cross(const Vector3& v) const
	{
		return Vector3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x);
	}
	void Cross(const Vector3& v, Vector3& result) const { result = Cross(v); }

	// XMVector3Normalize, SSE2 path: v / sqrt(dot); zero length -> 0, infinite length -> NaN.
	void Normalize()
	{
		const float lsq = Dot(*this);
		const float len = std::sqrt(lsq);
		if (lsq == 0.f) { x = y = z = 0.f; return; }
		if (std::isinf(lsq)) { x = y = z = std::numeric_limits<float>::quiet_NaN(); return; }
		x = x / len; y = y / len; z = z / len;
	}
	void Normalize(Vector3& result) const { result = *this; result.Normalize(); }

	static float Distance(const Vector3& a, const Vector3& b)
	{
		const Vector3 d(b.x - a.x, b.y - a.y, b.z - a.z);
		return d.Length();
	}
	static float DistanceSquared(const Vector3& a, const Vector3& b)
	{
		const Vector3 d(b.x - a.x, b.y - a.y, b.z - a.z);
		return d.LengthSquared();
	}
};

inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator*(const Vector3& a, const Vector3& b) { return Vector3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vector3 operator*(const Vector3& a, float s) { return Vector3(a.x * s, a.y * s, a.z * s); }
inline Vector3 operator*(float s, const Vector3& a) { return Vector3(a.x * s, a.y * s, a.z * s); }
inline Vector3 operator/(const Vector3& a, const Vector3& b) { return Vector3(a.x / b.x, a.y / b.y, a.z / b.z); }
// SimpleMath.inl:870-878: XMVectorScale(V, 1.f / S)
inline Vector3 operator/(const Vector3& a, float s) { const float r = 1.f / s; return Vector3(a.x * r, a.y * r, a.z * r); }

struct Plane : public XMFLOAT4
{
	Plane() : XMFLOAT4(0.f, 1.f, 0.f, 0.f) {}
	constexpr Plane(float ix, float iy, float iz, float iw) : XMFLOAT4(ix, iy, iz, iw) {}
	Plane(const Vector3& normal, float d) : XMFLOAT4(normal.x, normal.y, normal.z, d) {}
	// SimpleMath.inl:2773-2780 -> XMPlaneFromPoints
	Plane(const Vector3& p1, const Vector3& p2, const Vector3& p3)
	{
		Vector3 n = (p1 - p2).Cross(p1 - p3);
		n.Normalize();
		x = n.x; y = n.y; z = n.z; w = -n.Dot(p1);
	}
	// SimpleMath.inl:2782-2788 -> XMPlaneFromPointNormal (normal is NOT normalised)
	Plane(const Vector3& point, const Vector3& normal)
	{
		x = normal.x; y = normal.y; z = normal.z; w = -point.Dot(normal);
	}

	bool operator==(const Plane& p) const { return x == p.x && y == p.y && z == p.z && w == p.w; }
	bool operator!=(const Plane& p) const { return !(*this == p); }

	Vector3 Normal() const { return Vector3(x, y, z); }
	void Normal(const Vector3& n) { x = n.x; y = n.y; z = n.z; }
	float D() const { return w; }
	void D(float d) { w = d; }
};

struct Matrix
{
	float m[4][4];
};
} // namespace SimpleMath

// XMVector3TransformCoord(V, M): ((z*r2 + r3) + y*r1) + x*r0, then divide by w.
inline SimpleMath::Vector3 XMVector3TransformCoord(const SimpleMath::Vector3& v, const XMMATRIX& m)
{
	float o[4];
	for (int c = 0; c < 4; c++)
		o[c] = ((v.z * m.r[2][c] + m.r[3][c]) + v.y * m.r[1][c]) + v.x * m.r[0][c];
	return SimpleMath::Vector3(o[0] / o[3], o[1] / o[3], o[2] / o[3]);
}
} // namespace DirectX

using DirectX::XMFLOAT3;
using DirectX::XMFLOAT4;
using DirectX::XM_2PI;

// Skip the D3D12 body of Inc/Mesh.h (Mesh.h:1-2 guard), keep the POD it declares first.
#define MESH_H
struct VertexNormalColor
{
	DirectX::XMFLOAT3 Position;
	DirectX::XMFLOAT3 Normal;
	DirectX::XMFLOAT3 Color;

	explicit VertexNormalColor(const DirectX::XMFLOAT3 position = DirectX::XMFLOAT3(0, 0, 0),
							   const DirectX::XMFLOAT3 normal = DirectX::XMFLOAT3(0, 0, 0),
							   const DirectX::XMFLOAT3 color = DirectX::XMFLOAT3(0.25f, 0.25f, 0.25f))
		: Position(position), Normal(normal), Color(color)
	{
	}
};
