#include "Fracture.h"

#include "DT3D.h"
#include "Engine.h"

#include <algorithm>
#include <random>

namespace SurtrHost
{
std::vector<Vector3> GenerateSeeds(int seed, int cellCount)
{
	std::vector<Vector3> cellPointVec;
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> uniformDist(-0.5, 0.5);
	for (int i = 0; i < cellCount; i++)
	{
		const double x = uniformDist(gen);
		const double y = uniformDist(gen);
		const double z = uniformDist(gen);
		cellPointVec.emplace_back(x, y, z);
	}
	return cellPointVec;
}

std::vector<Vector3> GenerateRadialSeeds(int seed, int cellCount, double mean)
{
	std::vector<Vector3> cellPointVec;
	std::mt19937 gen(seed);
	std::uniform_real_distribution<double> directionUniformDist(-1.0, 1.0);
	std::exponential_distribution<double> lengthExpDist(1.0 / mean);
	for (int i = 0; i < cellCount; i++)
	{
		const double len = std::max(std::min(lengthExpDist(gen), 0.5), 1e-12);
		const double x = directionUniformDist(gen);
		const double y = directionUniformDist(gen);
		const double z = directionUniformDist(gen);
		Vector3 v(x, y, z);
		v.Normalize();
		v *= (float)len;
		cellPointVec.push_back(v);
	}
	return cellPointVec;
}

std::vector<VMACH::Polygon3D> GenerateVoronoi(const std::vector<Vector3>& cellPointVec) { return DT3D::VoronoiCells(cellPointVec); }

CompoundInfo ApplyFracture(const Compound& compound, const std::vector<VMACH::Polygon3D>& voroPolyVec)
{
	detail::FlatPolys pieces;
	for (const Piece* p : compound.PieceVec)
		pieces.add(p->Convex);
	detail::FlatCells cells;
	for (const VMACH::Polygon3D& cell : voroPolyVec)
		cells.add(cell);
	detail::Fragments fr;
	detail::run_event(pieces, cells, fr);

	CompoundInfo info;
	info.CompoundBind.push_back(std::set<int>());   // 0-th element is reserved (Surtr.cpp:2126)
	int current_cell = -1;
	for (size_t f = 0; f < fr.rec.size(); f++)
	{
		const surtr_fragment& r = fr.rec[f];
		const Poly::Polyhedron convex = fr.polyhedron(f);
		info.PieceVec.push_back(new Piece(convex, convex));   // mesh branch: "next" row f-1
		if ((int)r.cell != current_cell)   // fragments arrive cell-major: one bind set per non-empty cell (Surtr.cpp:2133-2146)
		{
			info.CompoundBind.push_back(std::set<int>());
			current_cell = (int)r.cell;
		}
		info.CompoundBind.back().insert((int)f);
		MassProperties m;
		m.Volume = r.volume;
		m.Centroid = Vector3(r.centroid[0], r.centroid[1], r.centroid[2]);
		std::copy(r.inertia, r.inertia + 6, m.Inertia);
		m.FaceCount = r.n_faces;
		info.PieceMass.push_back(m);
		info.PieceSourceCell.push_back((int)r.cell);
		info.PieceSourcePiece.push_back((int)r.piece);
	}
	return info;
}

void SetExtract(CompoundInfo& preResult)
{
	preResult.PieceExtractedConvex.resize(preResult.PieceVec.size(), nullptr);
	std::transform(preResult.PieceVec.begin(), preResult.PieceVec.end(), preResult.PieceExtractedConvex.begin(),
				   [](const Piece* p) { return Poly::ExtractFaces(p->Convex); });
}
} // namespace SurtrHost
