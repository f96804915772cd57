// SimpleMath.h -- the float32 value types of the cutting path for the headless host library.
//
// Same names and namespaces as the reference's math layer (DirectX::SimpleMath::Vector3 / Plane,
// ThirdParty/Inc/SimpleMath.h:631-681, SimpleMath.inl:729-1006, 2773-2788) so code written against the reference
// headers recompiles unchanged, with the DX12 / Windows dependencies gone.  Arithmetic contract: float32, every
// product and sum rounded separately, in the operand order of DirectXMath's SSE2 path --
// the same contract as surtr_b200/csrc/surtr_math.cuh on the device.  Build with -ffp-contract=off.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>

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

struct XMMATRIX   // row-major rows r[0..3]
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
	constexpr Vector3(float ix, float iy, float iz) : XMFLOAT3(ix, iy, iz) {}
	Vector3(double ix, double iy, double iz) : XMFLOAT3((float)ix, (float)iy, (float)iz) {}
	Vector3(int ix, int iy, int iz) : XMFLOAT3((float)ix, (float)iy, (float)iz) {}
	Vector3(const XMFLOAT3& v) : XMFLOAT3(v.x, v.y, v.z) {}

	bool operator==(const Vector3& v) const { return x == v.x && y == v.y && z == v.z; }
	bool operator!=(const Vector3& v) const { return !(*this == v); }

	Vector3& operator+=(const Vector3& v) { x = x + v.x; y = y + v.y; z = z + v.z; return *this; }
	Vector3& operator-=(const Vector3& v) { x = x - v.x; y = y - v.y; z = z - v.z; return *this; }
	Vector3& operator*=(const Vector3& v) { x = x * v.x; y = y * v.y; z = z * v.z; return *this; }
	Vector3& operator*=(float s) { x = x * s; y = y * s; z = z * s; return *this; }
	Vector3& operator/=(float s) { const float r = 1.f / s; x = x * r; y = y * r; z = z * r; return *this; }   // SimpleMath.inl:788-796
	Vector3 operator-() const { return Vector3(-x, -y, -z); }

	float Dot(const Vector3& v) const { return (x * v.x + y * v.y) + z * v.z; }   // SimpleMath.inl:918-925
	float LengthSquared() const { return Dot(*this); }
	float Length() const { return std::sqrt(Dot(*this)); }
	Vector3 Cross(const Vector3& v) const { return Vector3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x); }

	void Normalize()   // XMVector3Normalize: v / sqrt(dot); zero length -> 0, infinite -> NaN
	{
		const float lsq = Dot(*this);
		const float len = std::sqrt(lsq);
		if (lsq == 0.f) { x = y = z = 0.f; return; }
		if (std::isinf(lsq)) { x = y = z = std::numeric_limits<float>::quiet_NaN(); return; }
		x = x / len; y = y / len; z = z / len;
	}

	static float Distance(const Vector3& a, const Vector3& b) { return Vector3(b.x - a.x, b.y - a.y, b.z - a.z).Length(); }
};

inline Vector3 operator+(const Vector3& a, const Vector3& b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(const Vector3& a, const Vector3& b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator*(const Vector3& a, const Vector3& b) { return Vector3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vector3 operator*(const Vector3& a, float s) { return Vector3(a.x * s, a.y * s, a.z * s); }
inline Vector3 operator*(float s, const Vector3& a) { return Vector3(a.x * s, a.y * s, a.z * s); }
inline Vector3 operator/(const Vector3& a, float s) { const float r = 1.f / s; return Vector3(a.x * r, a.y * r, a.z * r); }   // SimpleMath.inl:870-878

struct Plane : public XMFLOAT4
{
	Plane() : XMFLOAT4(0.f, 1.f, 0.f, 0.f) {}
	constexpr Plane(float ix, float iy, float iz, float iw) : XMFLOAT4(ix, iy, iz, iw) {}
	Plane(const Vector3& normal, float d) : XMFLOAT4(normal.x, normal.y, normal.z, d) {}
	Plane(const Vector3& p1, const Vector3& p2, const Vector3& p3)   // SimpleMath.inl:2773-2780
	{
		Vector3 n = (p1 - p2).Cross(p1 - p3);
		n.Normalize();
		x = n.x; y = n.y; z = n.z; w = -n.Dot(p1);
	}
	Plane(const Vector3& point, const Vector3& normal)   // SimpleMath.inl:2782-2788, normal not normalised
	{
		x = normal.x; y = normal.y; z = normal.z; w = -point.Dot(normal);
	}
	bool operator==(const Plane& p) const { return x == p.x && y == p.y && z == p.z && w == p.w; }
	Vector3 Normal() const { return Vector3(x, y, z); }
	float D() const { return w; }
};
} // namespace SimpleMath

// XMVector3TransformCoord: ((z*r2 + r3) + y*r1) + x*r0, then divide by w (Poly.cpp:580-585)
inline SimpleMath::Vector3 XMVector3TransformCoord(const SimpleMath::Vector3& v, const XMMATRIX& m)
{
	float o[4];
	for (int c = 0; c < 4; c++)
		o[c] = ((v.z * m.r[2][c] + m.r[3][c]) + v.y * m.r[1][c]) + v.x * m.r[0][c];
	return SimpleMath::Vector3(o[0] / o[3], o[1] / o[3], o[2] / o[3]);
}
} // namespace DirectX
