// fracture_demo.cpp -- the reference's fracture flow (Surtr::PrepareFracture -> Surtr::DoFracture, Src/Surtr.cpp:1747-1959)
// written against the host-side mirror classes of surtr_b200/host/, i.e. what a maintainer's code looks like after
// switching: same types (Poly::Polyhedron, VMACH::Polygon3D, Piece, Compound), same calls, the cutting on the GPU.
//
//   build:  make -C examples            (needs surtr_b200/libsurtr_host.so + libsurtr_b200.so, built by __graft_entry__.build())
//   run:    examples/fracture_demo model.obj [scale] [seeds]      on a machine with a B200
//
// There is no CPU fallback: without a usable GPU the first GPU-backed call throws std::runtime_error.
#include "Fracture.h"

#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>

using DirectX::SimpleMath::Vector3;

// v / f lines only; x negated and winding flipped like Surtr::LoadModelData (Surtr.cpp:2683-2727)
static bool load_obj(const char* path, float scale, std::vector<Vector3>& vertices, std::vector<int>& indices)
{
	std::ifstream in(path);
	if (!in)
		return false;
	std::string line;
	while (std::getline(in, line))
	{
		std::istringstream ss(line);
		std::string tag;
		ss >> tag;
		if (tag == "v")
		{
			float x, y, z;
			ss >> x >> y >> z;
			vertices.emplace_back(-x * scale, y * scale, z * scale);
		}
		else if (tag == "f")
		{
			std::vector<int> f;
			std::string w;
			while (ss >> w)
				f.push_back(std::stoi(w.substr(0, w.find('/'))) - 1);
			for (size_t k = 1; k + 1 < f.size(); k++)
			{
				indices.push_back(f[0]);
				indices.push_back(f[k + 1]);
				indices.push_back(f[k]);
			}
		}
	}
	return !vertices.empty() && !indices.empty();
}

int main(int argc, char** argv)
{
	if (argc < 2)
	{
		std::fprintf(stderr, "usage: %s model.obj [scale=70] [seeds=32]\n", argv[0]);
		return 2;
	}
	const float scale = argc > 2 ? std::stof(argv[2]) : 70.f;
	const int n_seeds = argc > 3 ? std::stoi(argv[3]) : 32;
	std::vector<Vector3> vertices;
	std::vector<int> indices;
	if (!load_obj(argv[1], scale, vertices, indices))
	{
		std::fprintf(stderr, "cannot read %s\n", argv[1]);
		return 2;
	}
	try
	{
		SurtrHost::FractureArgs args;
		args.InitialDecomposeCellCnt = n_seeds;

		// Surtr::PrepareFracture: ACH + mesh polyhedron + initial decomposition + refit
		SurtrHost::PreparedObject obj = SurtrHost::PrepareFracture(vertices, indices, SurtrHost::GenerateSeeds(args.Seed, n_seeds), args);
		std::printf("PrepareFracture: ACH %zu vertices, mesh %zu vertices -> %zu pieces\n", obj.ACH.size(), obj.Mesh.size(), obj.Initial.PieceVec.size());

		// the compound PrepareFracture hands to InitCompound: all pieces, bind-set order (Surtr.cpp:1816-1826)
		SurtrHost::Compound compound;
		for (const std::set<int>& bind : obj.Initial.CompoundBind)
			for (const int i : bind)
			{
				compound.PieceVec.push_back(obj.Initial.PieceVec[i]);
				compound.PieceExtractedConvex.push_back(obj.Initial.PieceExtractedConvex[i]);
			}

		// Surtr::DoFracture at an impact point: a radial pattern, generated once, stays resident on the GPU
		SurtrHost::FractureStorage storage;
		storage.MaxAxisScale = obj.MaxAxisScale;
		storage.PartialFracturePattern = SurtrHost::GenerateFracturePattern(args.Seed, 32, args.PartialFracturePatternDist);
		storage.GeneralFracturePattern = SurtrHost::GenerateFracturePattern(args.Seed, 32, args.GeneralFracturePatternDist);
		std::vector<Vector3> sphere;   // unit-sphere samples scaled by 0.5 (m_spherePointCloud, Surtr.cpp:1506-1516)
		for (int i = 0; i < 42; i++)
		{
			const float z = 1.f - 2.f * (i + 0.5f) / 42.f, r = std::sqrt(1.f - z * z), phi = 2.399963f * i;
			sphere.emplace_back(0.5f * r * std::cos(phi), 0.5f * r * std::sin(phi), 0.5f * z);
		}
		args.ImpactPosition = DirectX::XMFLOAT3(vertices[0].x, vertices[0].y, vertices[0].z);
		args.ImpactRadius = 0.3f * obj.MaxAxisScale;
		for (const bool partial : { false, true })
		{
			args.PartialFracture = partial;
			SurtrHost::CompoundInfo info;
			const std::vector<SurtrHost::Compound> result = SurtrHost::DoFracture(compound, storage, sphere, args, &info);
			double volume = 0.0;
			for (const SurtrHost::MassProperties& m : info.PieceMass)
				volume += m.Volume;
			std::printf("DoFracture(%s): %zu pieces in %zu compounds, convex volume %.6f\n", partial ? "partial" : "general",
						info.PieceVec.size(), result.size(), volume);
			for (size_t b = 0; b < result.size() && b < 4; b++)
			{
				std::vector<SurtrHost::MassProperties> parts;
				for (const int i : info.CompoundBind[b])
					parts.push_back(info.PieceMass[i]);
				const SurtrHost::MassProperties m = SurtrHost::CombineMass(parts, 10.f);
				std::printf("  compound %zu: %zu pieces, mass %.4f, centre (%.3f %.3f %.3f), Ixx Iyy Izz %.4f %.4f %.4f\n", b, parts.size(), m.Volume,
							m.Centroid.x, m.Centroid.y, m.Centroid.z, m.Inertia[0], m.Inertia[1], m.Inertia[2]);
			}
		}
	}
	catch (const std::exception& e)
	{
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
