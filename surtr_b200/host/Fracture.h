// Fracture.h -- headless mirror of the fracture members of class Surtr (Inc/Surtr.h:89-155, 172-215).
//
// The reference keeps these as private members of its DX12 application class; here they are a plain namespace so
// the path runs without a window, a device or PhysX.  ApplyFracture is the drop-in for Surtr::ApplyFracture
// (Surtr.cpp:2098-2149): instead of one thread-pool task per cell (m_fractureTask, :1457-1504) it packs the
// compound's convex pieces and the cell planes once, runs ONE GPU fracture event through the C ABI and unpacks the
// fragments in the reference's order.  The mesh branch of m_fractureTask (:1470-1500) is a second event over the same
// resident cells (Piece::Mesh polyhedra, cut in the global-memory tier of K3) followed by the island split on the host.
#pragma once

#include <utility>
#include "Engine.h"
#include "Poly.h"
#include "VMACH.h"

#include <set>
#include <vector>

namespace SurtrHost
{
using DirectX::XMFLOAT3;
using DirectX::SimpleMath::Vector3;

struct FractureArgs   // Inc/Surtr.h:89-110, same defaults
{
	int ICHIncludePointLimit = 20;
	float ACHPlaneGapInverse = 2000.0f;
	int RefittingPointLimit = 4;
	int Seed = 46354;
	XMFLOAT3 ImpactPosition = XMFLOAT3(0.f, 0.f, 0.f);
	float ImpactRadius = 1.0f;
	bool RadialMode = true;
	bool PartialFracture = true;
	float PartialFracturePatternDist = 0.01f;
	float GeneralFracturePatternDist = 1.0f;
	int InitialDecomposeCellCnt = 64;
	int PartialFracturePatternCellCnt = 128;
	int GeneralFracturePatternCellCnt = 1024;
	float TargetAdder = 0.01f;
};

struct Piece   // Inc/Surtr.h:113-119; always allocated on the heap
{
	Poly::Polyhedron Convex;
	Poly::Polyhedron Mesh;
	Piece(const Poly::Polyhedron& convex, const Poly::Polyhedron& mesh) : Convex(convex), Mesh(mesh) {}
	Piece(Poly::Polyhedron&& convex, Poly::Polyhedron&& mesh) : Convex(std::move(convex)), Mesh(std::move(mesh)) {}   // (a Vertex owns a heap vector: no deep copies on the fracture path)
	Piece(const Poly::Polyhedron& convex, Poly::Polyhedron&& mesh) : Convex(convex), Mesh(std::move(mesh)) {}
};

typedef std::vector<std::vector<int>> Extract;

struct MassProperties   // per piece, unit density; replaces PxRigidBodyExt::updateMassAndInertia (Surtr.cpp:2520)
{
	double Volume = 0.0;
	Vector3 Centroid;
	float Inertia[6] = { 0, 0, 0, 0, 0, 0 };   // Ixx Iyy Izz Ixy Ixz Iyz about the centroid
	int FaceCount = 0;
};

struct CompoundInfo   // Inc/Surtr.h:123-128 (+ the per-piece results K4 computes)
{
	std::vector<Piece*> PieceVec;
	std::vector<Extract*> PieceExtractedConvex;
	std::vector<std::set<int>> CompoundBind;   // [0] reserved for "outside"; one set per non-empty cell, cell order
	std::vector<MassProperties> PieceMass;
	std::vector<int> PieceSourceCell, PieceSourcePiece;
};

struct Compound   // Inc/Surtr.h:130-134
{
	std::vector<Piece*> PieceVec;
	std::vector<Extract*> PieceExtractedConvex;
};

// Surtr::GenerateVoronoi(cellCount) seeds (Surtr.cpp:1984-2001): mt19937(seed), U(-0.5, 0.5), x/y/z order.
std::vector<Vector3> GenerateSeeds(int seed, int cellCount);
// Radial pattern seeds of Surtr::GenerateFracturePattern (Surtr.cpp:2072-2096).
std::vector<Vector3> GenerateRadialSeeds(int seed, int cellCount, double mean);
// Surtr::GenerateVoronoi(points): DT3D-derived cells (voro++ is not vendored).
std::vector<VMACH::Polygon3D> GenerateVoronoi(const std::vector<Vector3>& cellPointVec);

// Surtr::PrepareFracture (Surtr.cpp:1747-1827): ICH normals -> bounding box -> k-DOP with gap -> ACH (2x bbox clipped by
// the k-DOP) -> [mesh polyhedron from the triangle list, step 7] -> Voronoi cells of `seeds` scaled by the bbox extent
// and translated to its centre -> ApplyFracture on the single (ACH, mesh) piece -> Refitting -> SetExtract.
// The overload without indices runs the convex branch only (Piece::Mesh = Piece::Convex, no refit): the configuration
// BASELINE.json's throughput metric is quoted on.  The pattern caches (step 9) are GenerateRadialSeeds + GenerateVoronoi,
// built by the caller when it fractures again.
struct PreparedObject
{
	Vector3 BBCenter, MinBB, MaxBB;
	float MaxAxisScale = 0.f;
	int ICHFaceCnt = 0;
	Poly::Polyhedron ACH, Mesh;
	std::vector<VMACH::Polygon3D> Cells;
	CompoundInfo Initial;
};
PreparedObject PrepareFracture(const std::vector<Vector3>& vertices, const std::vector<Vector3>& seeds, const FractureArgs& args = FractureArgs());
PreparedObject PrepareFracture(const std::vector<Vector3>& vertices, const std::vector<int>& indices, const std::vector<Vector3>& seeds,
							   const FractureArgs& args = FractureArgs());

// Surtr::ApplyFracture (Surtr.cpp:2098-2149).  partial = true keeps the pieces that ConvexOutOfSphere puts outside the
// impact sphere uncut, first in PieceVec and bound in CompoundBind[0] (they are the caller's Piece objects, not copies).
// meshBranch = false skips the second clip and hands every piece its convex as mesh.
// Throws std::runtime_error on a C-ABI failure.
CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec,
						   const std::vector<Vector3>& spherePointCloud, bool partial, const FractureArgs& args = FractureArgs(),
						   bool meshBranch = true);
CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec, bool meshBranch = true);
// Surtr::ConvexOutOfSphere (Surtr.cpp:2415-2458), MergeOutOfImpact (:2368-2403), HandleConvexIsland (:2203-2366).
bool ConvexOutOfSphere(const Poly::Polyhedron& polyhedron, const Extract* extract, const std::vector<Vector3>& spherePointCloud,
					   const Vector3 origin, const float radius);
void MergeOutOfImpact(CompoundInfo& compoundInfo, const std::vector<Vector3>& spherePointCloud, const FractureArgs& args = FractureArgs());
void HandleConvexIsland(CompoundInfo& compoundInfo);
// Surtr::GenerateFracturePattern (Surtr.cpp:2072-2096): radial seeds -> cells in the unit box.
std::vector<VMACH::Polygon3D> GenerateFracturePattern(int seed, int cellCount, double mean);

struct FractureStorage   // the members of Inc/Surtr.h:136-155 that DoFracture reads
{
	float MaxAxisScale = 1.f;
	std::vector<VMACH::Polygon3D> PartialFracturePattern, GeneralFracturePattern;   // in the unit box, as generated

	// Flat copy of a pattern for the device (built on first use; it is uploaded once per context and then only placed).
	const detail::FlatPattern& Resident(bool partial) const;
	void PatternsChanged();   // call after editing a pattern

private:
	mutable detail::FlatPattern m_flat[2];
};
// Surtr::DoFracture (Surtr.cpp:1885-1959): place the pattern at the impact point (scale 2 x MaxAxisScale), ApplyFracture,
// SetExtract, [MergeOutOfImpact], HandleConvexIsland, Refitting, SetExtract -> one Compound per bind set (the reserved 0-th
// set included, even when empty, as in the reference).  The pieces are expected in world space (ExecuteFractureRoutine
// transforms them first, :1846-1852; use Poly::Transform).  out_info receives the flat piece list with its mass properties.
std::vector<Compound> DoFracture(const Compound& targetCompound, const FractureStorage& storage, const std::vector<Vector3>& spherePointCloud,
								 const FractureArgs& args = FractureArgs(), CompoundInfo* out_info = nullptr);
// Mass, centre of mass and inertia about it of a compound of pieces at uniform density (parallel-axis theorem): the
// quantities PxRigidBodyExt::updateMassAndInertia(body, 10.0f) derives at Surtr.cpp:2520.  Result.Volume holds the mass.
MassProperties CombineMass(const std::vector<MassProperties>& pieces, float density = 10.0f);
// Surtr::CheckMeshIsland (Surtr.cpp:2171-2199): connected components of the ring graph.
std::vector<std::set<int>> CheckMeshIsland(const Poly::Polyhedron& polyhedron);
// Surtr::SetExtract (Surtr.cpp:2151-2155).
void SetExtract(CompoundInfo& preResult);
// Surtr::Refitting (Surtr.cpp:2405-2413) = m_refittingTask (:1449-1455) for every piece, batched: the ICH normals of
// piece->Mesh (<= RefittingPointLimit points, host, VMACH::ConvexHull) -> k-DOP extents of piece->Mesh (one batched GPU
// call) -> piece->Convex clipped by its own [Min0, Max0, Min1, ...] plane list (one GPU event, one (piece, cell) pair per
// piece).  A piece whose convex is clipped away ends up with an empty Convex, as in the reference.
void Refitting(std::vector<Piece*>& targetPieceVec, const FractureArgs& args = FractureArgs(), std::vector<MassProperties>* mass = nullptr);
} // namespace SurtrHost
