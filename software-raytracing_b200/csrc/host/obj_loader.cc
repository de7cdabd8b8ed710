// obj_loader.cc -- see include/loader/obj_loader.h.
//
// Conversion rules taken from the reference (raylib/loader/obj_loader.cc):
//   :113     faces without a usable material get Lambertian(0.5)
//   :119-125 local bounds over ALL vertices of the file
//   :133-228 one StaticMesh per shape; per face three positions, optional texcoords (0,0 if absent), optional
//            normals (face normal if any corner lacks one); SetParameterization(u,v); CalculateBounds
//   :230-241 one shape -> the mesh is the root; several -> BVHNode(HitableList(meshes))
//   :342-398 material mapping: illum 4/6 with zero diffuse -> Dielectric(Ni, Tf); illum 3 -> Mirror(Kd);
//            otherwise MicrofacetMaterial (textures, Kd clamped to 0.95, roughness = Pr or
//            sqrt(2 / (Ns * mean(Ks) + 2)), metallic Pm, emissive Ke)
// Parsing follows tinyobjloader's observable behaviour for the statements listed in the header: 1-based and
// negative (relative) indices, shapes split at every `g` / `o`, polygons triangulated (here: as a fan), material
// defaults Kd = Ks = Ke = Tf = 0, Ns = 1, Ni = 1, illum = 0.
#include "loader/obj_loader.h"
#include "compact_mesh.h"
#include "core/int_types.h"
#include "core/logger.h"
#include "core/assertion.h"
#include "geom/primitives.h"
#include "geom/hit.h"
#include "render/material.h"
#include "render/image.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <sstream>

#define MAX_ALBEDO vec3(0.95f)

namespace
{
	struct RawMaterial
	{
		std::string name;
		float diffuse[3] = { 0, 0, 0 }, specular[3] = { 0, 0, 0 }, transmittance[3] = { 0, 0, 0 }, emission[3] = { 0, 0, 0 };
		float shininess = 1.0f, ior = 1.0f, roughness = 0.0f, metallic = 0.0f;
		int illum = 0;
		std::string diffuseTex, roughnessTex, metallicTex, emissiveTex, normalTex, bumpTex;
	};

	struct RawIndex { int v = -1, vt = -1, vn = -1; };
	struct RawShape
	{
		std::string name;
		std::vector<RawIndex> indices;       // three per triangle
		std::vector<int> materialIds;        // one per triangle
	};

	std::string Trim(const std::string& s)
	{
		size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
		return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
	}

	// Texture statements may carry options ("-bm 0.5 file.png"): the file name is the last token.
	std::string TextureName(const std::string& rest)
	{
		const std::string t = Trim(rest);
		if (t.empty() || t[0] != '-') return t;
		const size_t sp = t.find_last_of(" \t");
		return sp == std::string::npos ? t : t.substr(sp + 1);
	}

	void Parse3(std::istringstream& in, float* out) { in >> out[0]; if (!(in >> out[1])) { out[1] = out[2] = out[0]; return; } if (!(in >> out[2])) out[2] = out[1]; }

	// up to `count` blank-separated floats; missing ones keep their defaults (strtof rounds like the stream extraction did)
	void ParseFloats(const char* text, float* out, int count)
	{
		for (int i = 0; i < count; ++i)
		{
			char* end = nullptr;
			const float v = strtof(text, &end);
			if (end == text) return;
			out[i] = v;
			text = end;
		}
	}

	bool ParseMtl(const std::string& path, std::vector<RawMaterial>& out, std::map<std::string, int>& byName)
	{
		std::ifstream file(path);
		if (!file) return false;
		std::string line;
		RawMaterial* cur = nullptr;
		while (std::getline(file, line))
		{
			line = Trim(line);
			if (line.empty() || line[0] == '#') continue;
			std::istringstream in(line);
			std::string key;
			in >> key;
			std::string rest;
			std::getline(in, rest);
			std::istringstream args(rest);
			if (key == "newmtl") { out.push_back(RawMaterial()); cur = &out.back(); cur->name = Trim(rest); byName[cur->name] = (int)out.size() - 1; continue; }
			if (!cur) continue;
			if (key == "Kd") Parse3(args, cur->diffuse);
			else if (key == "Ks") Parse3(args, cur->specular);
			else if (key == "Ke") Parse3(args, cur->emission);
			else if (key == "Kt" || key == "Tf") Parse3(args, cur->transmittance);
			else if (key == "Ns") args >> cur->shininess;
			else if (key == "Ni") args >> cur->ior;
			else if (key == "illum") args >> cur->illum;
			else if (key == "Pr") args >> cur->roughness;
			else if (key == "Pm") args >> cur->metallic;
			else if (key == "map_Kd") cur->diffuseTex = TextureName(rest);
			else if (key == "map_Pr") cur->roughnessTex = TextureName(rest);
			else if (key == "map_Pm") cur->metallicTex = TextureName(rest);
			else if (key == "map_Ke") cur->emissiveTex = TextureName(rest);
			else if (key == "norm") cur->normalTex = TextureName(rest);
			else if (key == "map_bump" || key == "map_Bump" || key == "bump") cur->bumpTex = TextureName(rest);
		}
		return true;
	}

	int FixIndex(int idx, size_t count) { return idx > 0 ? idx - 1 : (idx < 0 ? (int)count + idx : -1); }

	// One face corner -- v, v/vt, v//vn or v/vt/vn -- at `c`, which is advanced past the token.  1-based and negative (relative) indices.
	bool ParseCornerInPlace(const char*& c, size_t nv, size_t nvt, size_t nvn, RawIndex& out)
	{
		char* end = nullptr;
		long parts[3] = { 0, 0, 0 };
		parts[0] = strtol(c, &end, 10);
		c = end;
		for (int part = 1; part < 3 && *c == '/'; ++part)
		{
			++c;
			if (*c != '/' && *c != ' ' && *c != '\t' && *c != '\0') { parts[part] = strtol(c, &end, 10); c = end; }
		}
		while (*c && *c != ' ' && *c != '\t') ++c;
		if (parts[0] == 0) return false;
		out.v = FixIndex((int)parts[0], nv);
		out.vt = parts[1] ? FixIndex((int)parts[1], nvt) : -1;
		out.vn = parts[2] ? FixIndex((int)parts[2], nvn) : -1;
		return out.v >= 0 && (size_t)out.v < nv;
	}

	float PhongSpecularToRoughness(const vec3& specularPower, float shininess)
	{
		const float intensity = (specularPower.x + specularPower.y + specularPower.z) / 3.0f;
		return std::sqrt(2.0f / (shininess * intensity + 2.0f));
	}
}

// Every mesh builds its own BVH (geom/static_mesh.cc:80-95) from a split-axis stream keyed by the node's position in its
// own tree, so the meshes of a model are independent: one worker per core instead of the reference's serial loop
// (loader/obj_loader.cc:31-37).  The result does not depend on the thread count.
void OBJModel::FinalizeAllMeshes()
{
	const unsigned threads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)staticMeshes.size()));
	if (threads <= 1 || staticMeshes.size() < 4) { for (StaticMesh* mesh : staticMeshes) mesh->Finalize(); return; }
	std::atomic<size_t> next(0);
	auto worker = [&]() { for (size_t i = next.fetch_add(1); i < staticMeshes.size(); i = next.fetch_add(1)) staticMeshes[i]->Finalize(); };
	std::vector<std::thread> pool;
	for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
	worker();
	for (std::thread& th : pool) th.join();
}

void OBJLoader::Initialize() { LOG("Initialize obj loader"); }
void OBJLoader::Destroy() { LOG("Destroy obj loader"); }

bool OBJLoader::LoadModelFromFile(const char* filepath, OBJModel* outModel)
{
	CHECK(outModel != nullptr);
	if (!outModel) return false;
	OBJLoader loader;
	return loader.LoadFromFile(filepath, *outModel);
}

OBJLoader::OBJLoader() {}
OBJLoader::~OBJLoader() {}

bool OBJLoader::LoadFromFile(const char* filepath, OBJModel& outModel)
{
	if (filepath == nullptr)
	{
		LOG("%s: filepath was null", __FUNCTION__);
		return false;
	}
	std::ifstream file(filepath);
	if (!file)
	{
		LOG("%s: cannot open: %s", __FUNCTION__, filepath);
		return false;
	}
	const std::string objpath(filepath);
	std::string basedir;
	if (objpath.find_last_of("/\\") != std::string::npos) basedir = objpath.substr(0, objpath.find_last_of("/\\") + 1);

	std::vector<float> positions, texcoords, normals;
	std::vector<RawMaterial> rawMaterials;
	std::map<std::string, int> materialByName;
	std::vector<RawShape> shapes;
	RawShape current;
	int currentMaterial = -1;
	int32 nonTriangleFaces = 0;

	// The whole file in memory; geometry statements (v / vt / vn / f: nearly every line of a large model) are parsed in
	// place with strtof / strtol -- no per-line string or stream objects --, everything else goes through a stream.
	std::string text((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
	text.push_back('\n');
	std::string line, key;
	std::vector<RawIndex> corners;
	for (size_t lineBegin = 0; lineBegin < text.size();)
	{
		size_t lineEnd = text.find('\n', lineBegin);
		if (lineEnd == std::string::npos) lineEnd = text.size();
		char* c = &text[lineBegin];
		text[lineEnd] = '\0';      // terminate the line in place for strtof / strtol
		lineBegin = lineEnd + 1;
		while (*c == ' ' || *c == '\t' || *c == '\r') ++c;
		if (*c == '\0' || *c == '#') continue;
		const bool blankAfter1 = c[1] == ' ' || c[1] == '\t', blankAfter2 = c[1] != '\0' && (c[2] == ' ' || c[2] == '\t');
		if (c[0] == 'v' && blankAfter1) { float p[3] = { 0, 0, 0 }; ParseFloats(c + 1, p, 3); positions.insert(positions.end(), p, p + 3); continue; }
		if (c[0] == 'v' && c[1] == 't' && blankAfter2) { float t[2] = { 0, 0 }; ParseFloats(c + 2, t, 2); texcoords.insert(texcoords.end(), t, t + 2); continue; }
		if (c[0] == 'v' && c[1] == 'n' && blankAfter2) { float n[3] = { 0, 0, 0 }; ParseFloats(c + 2, n, 3); normals.insert(normals.end(), n, n + 3); continue; }
		if (c[0] == 'f' && blankAfter1)
		{
			corners.clear();
			bool ok = true;
			const char* q = c + 1;
			for (;;)
			{
				while (*q == ' ' || *q == '\t' || *q == '\r') ++q;
				if (!*q) break;
				RawIndex idx;
				if (!ParseCornerInPlace(q, positions.size() / 3, texcoords.size() / 2, normals.size() / 3, idx)) { ok = false; break; }
				corners.push_back(idx);
			}
			if (!ok || corners.size() < 3) continue;
			if (corners.size() > 3) ++nonTriangleFaces;
			for (size_t i = 1; i + 1 < corners.size(); ++i)
			{
				current.indices.push_back(corners[0]); current.indices.push_back(corners[i]); current.indices.push_back(corners[i + 1]);
				current.materialIds.push_back(currentMaterial);
			}
			continue;
		}
		line = Trim(std::string(c));
		std::istringstream in(line);
		in >> key;
		if (false) {}
		else if (key == "g" || key == "o")
		{
			if (!current.materialIds.empty()) shapes.push_back(std::move(current));
			current = RawShape();
			std::string rest;
			std::getline(in, rest);
			current.name = Trim(rest);
		}
		else if (key == "usemtl")
		{
			std::string rest;
			std::getline(in, rest);
			auto it = materialByName.find(Trim(rest));
			currentMaterial = it != materialByName.end() ? it->second : -1;
		}
		else if (key == "mtllib")
		{
			std::string rest;
			std::getline(in, rest);
			std::istringstream names(rest);
			std::string name;
			while (names >> name)
				if (ParseMtl(basedir + name, rawMaterials, materialByName)) break;
		}
	}
	if (!current.materialIds.empty()) shapes.push_back(std::move(current));

	if (shapes.empty())
	{
		LOG("%s: No shapes found in: %s", __FUNCTION__, filepath);
		return false;
	}
	LOG("%s: Load %s", __FUNCTION__, filepath);
	LOG("\tTotal shapes: %d", (int32)shapes.size());
	LOG("\tTotal vertices: %d", (int32)(positions.size() / 3));
	LOG("\tTotal materials: %d", (int32)rawMaterials.size());
	if (nonTriangleFaces > 0) LOG("\t%d polygons were triangulated as fans", nonTriangleFaces);

	Lambertian* const fallbackMaterial = new Lambertian(vec3(0.5f, 0.5f, 0.5f));

	// ---- images (obj_loader.cc:259-292) and materials (:294-400) ----
	auto preload = [&](const std::string& name) {
		if (name.empty() || imageDB.find(name) != imageDB.end() || basedir.empty()) return;
		Image2D* image = ImageIO::LoadImage2DFromFile((basedir + name).c_str());
		imageDB.insert(std::make_pair(name, std::shared_ptr<Image2D>(image)));
	};
	for (const RawMaterial& m : rawMaterials)
	{
		preload(m.diffuseTex); preload(m.roughnessTex); preload(m.metallicTex); preload(m.emissiveTex); preload(m.normalTex); preload(m.bumpTex);
	}
	LOG("\t%u image files has been loaded", (uint32)imageDB.size());
	auto findImage = [&](const std::string& name) {
		auto it = imageDB.find(name);
		return it != imageDB.end() ? it->second : std::shared_ptr<Image2D>();
	};
	materials.assign(rawMaterials.size(), nullptr);
	for (size_t i = 0; i < rawMaterials.size(); ++i)
	{
		const RawMaterial& raw = rawMaterials[i];
		std::shared_ptr<Image2D> albedoImage = findImage(raw.diffuseTex), roughnessImage = findImage(raw.roughnessTex);
		std::shared_ptr<Image2D> metallicImage = findImage(raw.metallicTex), emissiveImage = findImage(raw.emissiveTex);
		std::shared_ptr<Image2D> normalImage = findImage(raw.normalTex);
		if (normalImage == nullptr) normalImage = findImage(raw.bumpTex);

		const vec3 albedoConstant = min(MAX_ALBEDO, vec3(raw.diffuse[0], raw.diffuse[1], raw.diffuse[2]));
		const bool transparentIllum = raw.illum == 4 || raw.illum == 6;
		const bool zeroDiffuse = raw.diffuseTex.empty() && albedoConstant == vec3(0.0f);
		if (transparentIllum && zeroDiffuse)
			materials[i] = new Dielectric(raw.ior, vec3(raw.transmittance[0], raw.transmittance[1], raw.transmittance[2]));
		else if (raw.illum == 3)
			materials[i] = new Mirror(albedoConstant);
		else
		{
			MicrofacetMaterial* M = new MicrofacetMaterial;
			if (albedoImage) M->SetAlbedoTexture(albedoImage);
			if (normalImage) M->SetNormalTexture(normalImage);
			if (roughnessImage) M->SetRoughnessTexture(roughnessImage);
			if (metallicImage) M->SetMetallicTexture(metallicImage);
			if (emissiveImage) M->SetEmissiveTexture(emissiveImage);
			M->SetAlbedoFallback(albedoConstant);
			if (raw.roughness > 0.0f) M->SetRoughnessFallback(raw.roughness);
			else M->SetRoughnessFallback(PhongSpecularToRoughness(vec3(raw.specular[0], raw.specular[1], raw.specular[2]), raw.shininess));
			M->SetMetallicFallback(raw.metallic);
			M->SetEmissiveFallback(vec3(raw.emission[0], raw.emission[1], raw.emission[2]));
			materials[i] = M;
		}
	}

	vec3 localMinBound(FLOAT_MAX, FLOAT_MAX, FLOAT_MAX), localMaxBound(-FLOAT_MAX, -FLOAT_MAX, -FLOAT_MAX);
	for (size_t i = 0; i + 2 < positions.size(); i += 3)
	{
		const vec3 v(positions[i], positions[i + 1], positions[i + 2]);
		localMinBound = min(localMinBound, v);
		localMaxBound = max(localMaxBound, v);
	}

	// ---- shapes -> StaticMesh (obj_loader.cc:133-228) ----
	int32 numInvalidTexcoords = 0;
	for (const RawShape& shape : shapes)
	{
		StaticMesh* mesh = new StaticMesh;
		// fast path (compact_mesh.h): the faces stay plain arrays next to the mesh; RAYLIB_B200_OBJ_OBJECTS=1 builds the
		// reference's Triangle objects instead (the two paths flatten to the same bits, tests/test_cpu_host.py)
		const char* objectsEnv = getenv("RAYLIB_B200_OBJ_OBJECTS");
		const bool wantObjects = objectsEnv && atoi(objectsEnv) != 0;
		RtCompactMesh* compact = wantObjects ? nullptr : RtAttachCompactMesh(mesh);
		if (compact)
		{
			const size_t faces = shape.materialIds.size();
			compact->positions.reserve(3 * faces); compact->normals.reserve(3 * faces); compact->texcoords.reserve(6 * faces); compact->materials.reserve(faces);
		}
		for (size_t face = 0; face < shape.materialIds.size(); ++face)
		{
			vec3 p[3], n[3];
			float us[3], vs[3];
			bool validNormal = true;
			for (int c = 0; c < 3; ++c)
			{
				const RawIndex& idx = shape.indices[3 * face + c];
				p[c] = vec3(positions[3 * idx.v], positions[3 * idx.v + 1], positions[3 * idx.v + 2]);
				us[c] = vs[c] = 0.0f;
				if (idx.vt >= 0 && (size_t)idx.vt < texcoords.size() / 2) { us[c] = texcoords[2 * idx.vt]; vs[c] = texcoords[2 * idx.vt + 1]; }
				else ++numInvalidTexcoords;
				n[c] = vec3(0.0f, 0.0f, 0.0f);
				if (idx.vn >= 0 && (size_t)idx.vn < normals.size() / 3) n[c] = vec3(normals[3 * idx.vn], normals[3 * idx.vn + 1], normals[3 * idx.vn + 2]);
				else validNormal = false;
			}
			if (!validNormal)
			{
				const vec3 faceNormal = cross(p[1] - p[0], p[2] - p[0]);
				n[0] = n[1] = n[2] = normalize(faceNormal);
			}
			Material* faceMaterial = fallbackMaterial;
			const int mid = shape.materialIds[face];
			if (0 <= mid && mid < (int)materials.size() && materials[mid] != nullptr) faceMaterial = materials[mid];
			if (compact)
			{
				for (int c = 0; c < 3; ++c) { compact->positions.push_back(p[c]); compact->normals.push_back(n[c]); compact->texcoords.push_back(us[c]); compact->texcoords.push_back(vs[c]); }
				compact->materials.push_back(faceMaterial);
				continue;
			}
			Triangle T(p[0], p[1], p[2], n[0], n[1], n[2], faceMaterial);
			T.SetParameterization(us[0], vs[0], us[1], vs[1], us[2], vs[2]);
			mesh->AddTriangle(T);
		}
		mesh->CalculateBounds();
		outModel.staticMeshes.push_back(mesh);
	}

	if (outModel.staticMeshes.size() == 1) outModel.rootObject = outModel.staticMeshes[0];
	else
	{
		std::vector<Hitable*> hitables;
		for (StaticMesh* mesh : outModel.staticMeshes) hitables.push_back(mesh);
		outModel.rootObject = new BVHNode(new HitableList(hitables), 0.0f, 0.0f);
	}
	outModel.localMinBound = localMinBound;
	outModel.localMaxBound = localMaxBound;
	if (numInvalidTexcoords > 0) LOG("WARNING: Num triangles with invalid UVs: %d", numInvalidTexcoords);
	LOG("> OBJ loading done");
	return true;
}
