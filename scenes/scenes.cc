// scenes/scenes.cc -- procedural benchmark / parity scenes, written as a CLIENT of the raylib API.
//
// This file uses nothing but the public raylib interface (the C entry points of raylib.h plus the
// exported C++ scene classes, exactly like the reference's src/main.cc:570-984 does) and is compiled
// twice from the same source:
//   * against include/ + libraylib_b200.so            -> scenes/lib/libscenes_b200.so   (the product)
//   * against /root/reference/raylib + libraylib_ref.so -> oracle/_ref/libscenes_ref.so   (the oracle)
// which is the strongest drop-in check this repository has.  The scenes are the five BASELINE.json
// configurations.  All randomness comes from the local generator below, so both
// builds construct bit-identical geometry.
#include "raylib.h"
#include "geom/sphere.h"
#include "geom/triangle.h"
#include "geom/static_mesh.h"
#include "geom/cube.h"
#include "geom/bvh.h"
#include "render/material.h"
#include "render/image.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

extern "C" void scene_hook_before_bvh_build(void);   // oracle build: re-keys the reference's RNG; product build: no-op
extern "C" void scene_hook_on_destroy(SceneHandle);  // oracle build: drops the driver's per-scene caches; product build: no-op

namespace
{
	struct Gen   // splitmix64
	{
		uint64_t s;
		explicit Gen(uint64_t seed) : s(seed) {}
		uint64_t next() { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
		float uni() { return (float)(next() >> 40) * (1.0f / 16777216.0f); }
		float range(float a, float b) { return a + (b - a) * uni(); }
	};

	// triangles handed to StaticMesh::AddTriangle while a demo scene is built (DemoSceneInfo::numTriangles counts them
// together with top-level Triangle elements, so both libraries report the scene's real size)
static uint64_t g_meshTriangles = 0;
static inline void countedAdd(StaticMesh* mesh, const Triangle& t) { mesh->AddTriangle(t); g_meshTriangles++; }

struct Owned
	{
		std::vector<Hitable*> hitables;
		std::vector<Material*> materials;
		std::vector<std::shared_ptr<Image2D>> images;
		ImageHandle sky = 0;
		~Owned()
		{
			for (Hitable* h : hitables) delete h;
			// Material has no virtual destructor in the reference ABI; these are PODs + texture pointers.
			for (Material* m : materials) ::operator delete(m);
			if (sky) Raylib_DestroyImage(sky);
		}
		template<typename T> T* keep(T* h) { hitables.push_back(h); return h; }
		template<typename T> T* mat(T* m) { materials.push_back(m); return m; }
	};

	std::mutex g_mutex;
	std::map<SceneHandle, Owned*> g_owned;

	void finalizeMesh(StaticMesh* mesh) { scene_hook_before_bvh_build(); mesh->Finalize(); }
	void finalizeScene(SceneHandle scene) { scene_hook_before_bvh_build(); Raylib_FinalizeScene(scene); }

	ImageHandle makeGradientSky(uint32_t w, uint32_t h, const vec3& horizon, const vec3& zenith)
	{
		ImageHandle handle = Raylib_CreateImage(w, h);
		Image2D* img = (Image2D*)handle;
		for (uint32_t y = 0; y < h; ++y)
		{
			const float t = (h > 1) ? (float)y / (float)(h - 1) : 0.0f;
			const vec3 c = (1.0f - t) * horizon + t * zenith;
			for (uint32_t x = 0; x < w; ++x) img->SetPixel((int32)x, (int32)y, Pixel(c.x, c.y, c.z, 1.0f));
		}
		return handle;
	}

	void addQuad(StaticMesh* mesh, const vec3& a, const vec3& b, const vec3& c, const vec3& d, const vec3& n, Material* m)
	{
		Triangle t0(a, b, c, n, n, n, m); t0.SetParameterization(0.0f, 0.0f, 1.0f, 0.0f, 1.0f, 1.0f);
		Triangle t1(a, c, d, n, n, n, m); t1.SetParameterization(0.0f, 0.0f, 1.0f, 1.0f, 0.0f, 1.0f);
		countedAdd(mesh, t0);
		countedAdd(mesh, t1);
	}

	// axis-aligned box rotated about Y by `deg`, normals pointing outwards; 5 faces (no bottom)
	void addBox(StaticMesh* mesh, const vec3& center, const vec3& half, float deg, Material* m)
	{
		const float rad = deg * 3.14159265f / 180.0f, cs = cosf(rad), sn = sinf(rad);
		auto P = [&](float x, float y, float z) { return vec3(center.x + cs * x + sn * z, center.y + y, center.z - sn * x + cs * z); };
		auto N = [&](float x, float y, float z) { return vec3(cs * x + sn * z, y, -sn * x + cs * z); };
		const float hx = half.x, hy = half.y, hz = half.z;
		addQuad(mesh, P(-hx, hy, -hz), P(-hx, hy, hz), P(hx, hy, hz), P(hx, hy, -hz), N(0, 1, 0), m);          // top
		addQuad(mesh, P(-hx, -hy, hz), P(hx, -hy, hz), P(hx, hy, hz), P(-hx, hy, hz), N(0, 0, 1), m);          // front
		addQuad(mesh, P(hx, -hy, -hz), P(-hx, -hy, -hz), P(-hx, hy, -hz), P(hx, hy, -hz), N(0, 0, -1), m);     // back
		addQuad(mesh, P(-hx, -hy, -hz), P(-hx, -hy, hz), P(-hx, hy, hz), P(-hx, hy, -hz), N(-1, 0, 0), m);     // left
		addQuad(mesh, P(hx, -hy, hz), P(hx, -hy, -hz), P(hx, hy, -hz), P(hx, hy, hz), N(1, 0, 0), m);          // right
	}

	// ---- icosphere --------------------------------------------------------------------------------
	struct IcoMesh { std::vector<vec3> verts; std::vector<uint32_t> idx; };

	IcoMesh makeIcosphere(int level)
	{
		IcoMesh m;
		const float t = (1.0f + sqrtf(5.0f)) * 0.5f;
		const vec3 base[12] = {
			{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
			{0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1} };
		for (const vec3& v : base) m.verts.push_back(normalize(v));
		const uint32_t faces[60] = {
			0,11,5, 0,5,1, 0,1,7, 0,7,10, 0,10,11, 1,5,9, 5,11,4, 11,10,2, 10,7,6, 7,1,8,
			3,9,4, 3,4,2, 3,2,6, 3,6,8, 3,8,9, 4,9,5, 2,4,11, 6,2,10, 8,6,7, 9,8,1 };
		m.idx.assign(faces, faces + 60);
		for (int l = 0; l < level; ++l)
		{
			std::map<uint64_t, uint32_t> mid;
			auto midpoint = [&](uint32_t a, uint32_t b) {
				const uint64_t key = a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
				auto it = mid.find(key);
				if (it != mid.end()) return it->second;
				m.verts.push_back(normalize((m.verts[a] + m.verts[b]) * 0.5f));
				const uint32_t id = (uint32_t)m.verts.size() - 1;
				mid[key] = id;
				return id;
			};
			std::vector<uint32_t> next;
			for (size_t f = 0; f < m.idx.size(); f += 3)
			{
				const uint32_t a = m.idx[f], b = m.idx[f + 1], c = m.idx[f + 2];
				const uint32_t ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
				const uint32_t tri[12] = { a, ab, ca, b, bc, ab, c, ca, bc, ab, bc, ca };
				next.insert(next.end(), tri, tri + 12);
			}
			m.idx.swap(next);
		}
		return m;
	}

	// smooth-shaded blob: icosphere with a low-frequency radial bump, scaled/translated
	StaticMesh* makeBlob(const IcoMesh& ico, const vec3& center, const vec3& scale, float bump, float phase, Material* m, bool withUV)
	{
		StaticMesh* mesh = new StaticMesh;
		std::vector<vec3> pos(ico.verts.size()), nrm(ico.verts.size());
		for (size_t i = 0; i < ico.verts.size(); ++i)
		{
			const vec3& d = ico.verts[i];
			const float r = 1.0f + bump * sinf(5.0f * d.x + phase) * sinf(4.0f * d.y + 1.3f * phase) * sinf(6.0f * d.z);
			pos[i] = center + scale * (d * r);
			nrm[i] = normalize(d / scale);
		}
		for (size_t f = 0; f < ico.idx.size(); f += 3)
		{
			const uint32_t a = ico.idx[f], b = ico.idx[f + 1], c = ico.idx[f + 2];
			Triangle t(pos[a], pos[b], pos[c], nrm[a], nrm[b], nrm[c], m);
			if (withUV)
			{
				auto U = [&](uint32_t i) { return 0.5f + atan2f(ico.verts[i].z, ico.verts[i].x) * 0.15915494f; };
				auto V = [&](uint32_t i) { return 0.5f + asinf(ico.verts[i].y) * 0.31830989f; };
				t.SetParameterization(2.0f * U(a), 2.0f * V(a), 2.0f * U(b), 2.0f * V(b), 2.0f * U(c), 2.0f * V(c));
			}
			countedAdd(mesh, t);
		}
		return mesh;
	}

	float terrainHeight(float x, float z) { return 0.3f * sinf(3.0f * x) * cosf(3.0f * z) + 0.05f * sinf(40.0f * x + 17.0f * z); }

	// G x G cells over [-extent, extent]^2, two triangles per cell
	StaticMesh* makeDisplacedGrid(int G, float extent, float yOffset, Material* m, bool flatNormals, float uvTiles)
	{
		StaticMesh* mesh = new StaticMesh;
		const float step = 2.0f * extent / (float)G;
		auto P = [&](int i, int j) { const float x = -extent + step * (float)i, z = -extent + step * (float)j; return vec3(x, yOffset + terrainHeight(x, z), z); };
		auto Nrm = [&](int i, int j) {
			const vec3 dx = P(i + 1, j) - P(i - 1, j), dz = P(i, j + 1) - P(i, j - 1);
			return normalize(cross(dz, dx));
		};
		for (int j = 0; j < G; ++j)
			for (int i = 0; i < G; ++i)
			{
				const vec3 p00 = P(i, j), p10 = P(i + 1, j), p01 = P(i, j + 1), p11 = P(i + 1, j + 1);
				const float u0 = uvTiles * (float)i / (float)G, u1 = uvTiles * (float)(i + 1) / (float)G;
				const float v0 = uvTiles * (float)j / (float)G, v1 = uvTiles * (float)(j + 1) / (float)G;
				if (flatNormals)
				{
					const vec3 na = normalize(cross(p01 - p00, p11 - p00)), nb = normalize(cross(p11 - p00, p10 - p00));
					Triangle a(p00, p01, p11, na, na, na, m); a.SetParameterization(u0, v0, u0, v1, u1, v1);
					Triangle b(p00, p11, p10, nb, nb, nb, m); b.SetParameterization(u0, v0, u1, v1, u1, v0);
					countedAdd(mesh, a); countedAdd(mesh, b);
				}
				else
				{
					const vec3 n00 = Nrm(i, j), n10 = Nrm(i + 1, j), n01 = Nrm(i, j + 1), n11 = Nrm(i + 1, j + 1);
					Triangle a(p00, p01, p11, n00, n01, n11, m); a.SetParameterization(u0, v0, u0, v1, u1, v1);
					Triangle b(p00, p11, p10, n00, n11, n10, m); b.SetParameterization(u0, v0, u1, v1, u1, v0);
					countedAdd(mesh, a); countedAdd(mesh, b);
				}
			}
		return mesh;
	}

	// ---- config 1: RandomSpheres (src/main.cc:913-958), scene seed 42 ---------------------------------
	void buildRandomSpheres(SceneHandle scene, Owned& own)
	{
		Gen g(42);
		auto add = [&](const vec3& c, float r, Material* m) { Raylib_AddSceneElement(scene, (SceneElementHandle)own.keep(new Sphere(c, r, m))); };
		add(vec3(0.0f, -1000.0f, 0.0f), 1000.0f, own.mat(new Lambertian(vec3(0.5f, 0.5f, 0.5f))));
		for (int a = -6; a < 6; ++a)
			for (int b = -6; b < 6; ++b)
			{
				const float choose = g.uni();
				const float cx = (float)a + 0.9f * g.uni();
				const float cz = (float)b + 0.9f * g.uni();
				const vec3 center(cx, 0.2f, cz);
				if ((center - vec3(4.0f, 0.2f, 0.0f)).Length() > 2.0f)
				{
					if (choose < 0.8f)
					{
						const float r0 = g.uni(), r1 = g.uni(), r2 = g.uni(), r3 = g.uni(), r4 = g.uni(), r5 = g.uni();
						add(center, 0.2f, own.mat(new Lambertian(vec3(r0 * r1, r2 * r3, r4 * r5))));
					}
					else if (choose < 0.95f)
					{
						const float r0 = g.uni(), r1 = g.uni(), r2 = g.uni(), r3 = g.uni();
						add(center, 0.2f, own.mat(new Metal(vec3(0.5f * (1.0f + r0), 0.5f * (1.0f + r1), 0.5f * (1.0f + r2)), 0.5f * r3)));
					}
					else add(center, 0.2f, own.mat(new Dielectric(1.5f)));
				}
			}
		add(vec3(0.0f, 1.0f, 0.0f), 1.0f, own.mat(new Dielectric(1.5f)));
		add(vec3(-2.0f, 1.0f, 0.0f), 1.0f, own.mat(new Lambertian(vec3(0.4f, 0.2f, 0.1f))));
		add(vec3(2.0f, 1.0f, 0.0f), 1.0f, own.mat(new Metal(vec3(0.7f, 0.6f, 0.5f), 0.0f)));
		own.sky = makeGradientSky(512, 256, vec3(1.0f, 1.0f, 1.0f), vec3(0.5f, 0.7f, 1.0f));
		Raylib_SetSkyPanorama(scene, own.sky);
		Raylib_SetSunIlluminance(scene, 0.0f, 0.0f, 0.0f);
	}

	// ---- config 2: procedural Cornell box, 36 triangles in one StaticMesh ------------------------------
	void buildCornell(SceneHandle scene, Owned& own)
	{
		Material* white = own.mat(new Lambertian(vec3(0.73f, 0.73f, 0.73f)));
		Material* red = own.mat(new Lambertian(vec3(0.65f, 0.05f, 0.05f)));
		Material* green = own.mat(new Lambertian(vec3(0.12f, 0.45f, 0.15f)));
		Material* light = own.mat(new DiffuseLight(vec3(15.0f, 15.0f, 15.0f)));
		StaticMesh* mesh = own.keep(new StaticMesh);
		addQuad(mesh, vec3(-1, 0, 1), vec3(1, 0, 1), vec3(1, 0, -1), vec3(-1, 0, -1), vec3(0, 1, 0), white);     // floor
		addQuad(mesh, vec3(-1, 2, -1), vec3(1, 2, -1), vec3(1, 2, 1), vec3(-1, 2, 1), vec3(0, -1, 0), white);    // ceiling
		addQuad(mesh, vec3(-1, 0, -1), vec3(1, 0, -1), vec3(1, 2, -1), vec3(-1, 2, -1), vec3(0, 0, 1), white);   // back
		addQuad(mesh, vec3(-1, 0, 1), vec3(-1, 0, -1), vec3(-1, 2, -1), vec3(-1, 2, 1), vec3(1, 0, 0), red);     // left
		addQuad(mesh, vec3(1, 0, -1), vec3(1, 0, 1), vec3(1, 2, 1), vec3(1, 2, -1), vec3(-1, 0, 0), green);      // right
		addQuad(mesh, vec3(-0.35f, 1.98f, -0.3f), vec3(0.35f, 1.98f, -0.3f), vec3(0.35f, 1.98f, 0.3f), vec3(-0.35f, 1.98f, 0.3f), vec3(0, -1, 0), light);
		addBox(mesh, vec3(0.35f, 0.3f, 0.35f), vec3(0.3f, 0.3f, 0.3f), -18.0f, white);      // short box
		addBox(mesh, vec3(-0.35f, 0.6f, -0.3f), vec3(0.3f, 0.6f, 0.3f), 20.0f, white);      // tall box
		addBox(mesh, vec3(0.0f, 1.5f, -0.7f), vec3(0.15f, 0.05f, 0.15f), 45.0f, white);     // small shelf (42 triangles in total)
		finalizeMesh(mesh);
		Raylib_AddSceneElement(scene, (SceneElementHandle)mesh);
		Raylib_SetSkyPanorama(scene, 0);
		Raylib_SetSunIlluminance(scene, 0.0f, 0.0f, 0.0f);
	}

	// ---- config 3: 2*G*G-triangle displaced grid, flat normals, white sky ("primary + AO") -------------
	void buildDisplacedGrid(SceneHandle scene, Owned& own, int G)
	{
		Material* white = own.mat(new Lambertian(vec3(1.0f, 1.0f, 1.0f)));
		StaticMesh* mesh = own.keep(makeDisplacedGrid(G, 5.0f, 0.0f, white, true, 1.0f));
		finalizeMesh(mesh);
		Raylib_AddSceneElement(scene, (SceneElementHandle)mesh);
		own.sky = makeGradientSky(1, 1, vec3(1.0f, 1.0f, 1.0f), vec3(1.0f, 1.0f, 1.0f));
		Raylib_SetSkyPanorama(scene, own.sky);
		Raylib_SetSunIlluminance(scene, 0.0f, 0.0f, 0.0f);
	}

	// ---- config 4: instance scatter, two-level BVH like the San Miguel OBJ path --------------------------
	void buildInstanceScatter(SceneHandle scene, Owned& own, int numInstances, int icoLevel)
	{
		Gen g(7);
		Material* mats[8] = {
			own.mat(new Lambertian(vec3(0.75f, 0.25f, 0.2f))), own.mat(new Lambertian(vec3(0.25f, 0.6f, 0.3f))),
			own.mat(new Lambertian(vec3(0.3f, 0.35f, 0.8f))), own.mat(new Lambertian(vec3(0.8f, 0.75f, 0.6f))),
			own.mat(new Metal(vec3(0.9f, 0.85f, 0.7f), 0.05f)), own.mat(new Metal(vec3(0.7f, 0.75f, 0.8f), 0.3f)),
			own.mat(new Dielectric(1.5f)),
			own.mat(MicrofacetMaterial::FromConstants(vec3(0.8f, 0.5f, 0.2f), 0.3f, 0.0f, vec3(0.0f))) };
		Material* groundMat = own.mat(new Lambertian(vec3(0.45f, 0.45f, 0.45f)));
		const IcoMesh ico = makeIcosphere(icoLevel);
		int side = 1; while (side * side < numInstances) ++side;
		const float spacing = 1.0f, half = 0.5f * spacing * (float)side;
		for (int k = 0; k < numInstances; ++k)
		{
			const int ix = k % side, iz = k / side;
			const float jx = g.range(-0.25f, 0.25f), jz = g.range(-0.25f, 0.25f);
			const float s = g.range(0.22f, 0.42f), sy = s * g.range(0.8f, 1.6f);
			const vec3 center(-half + spacing * ((float)ix + 0.5f) + jx, sy * 0.9f, -half + spacing * ((float)iz + 0.5f) + jz);
			StaticMesh* mesh = own.keep(makeBlob(ico, center, vec3(s, sy, s), 0.12f, g.range(0.0f, 6.28f), mats[k % 8], false));
			finalizeMesh(mesh);
			Raylib_AddSceneElement(scene, (SceneElementHandle)mesh);
		}
		StaticMesh* ground = own.keep(new StaticMesh);
		const float e = half + 5.0f;
		addQuad(ground, vec3(-e, 0, e), vec3(e, 0, e), vec3(e, 0, -e), vec3(-e, 0, -e), vec3(0, 1, 0), groundMat);
		finalizeMesh(ground);
		Raylib_AddSceneElement(scene, (SceneElementHandle)ground);
		own.sky = makeGradientSky(512, 256, vec3(0.9f, 0.9f, 0.95f), vec3(0.35f, 0.55f, 0.95f));
		Raylib_SetSkyPanorama(scene, own.sky);
		Raylib_SetSunIlluminance(scene, 20.0f, 20.0f, 20.0f);
		Raylib_SetSunDirection(scene, 0.0f, -1.0f, -0.5f);
	}

	// ---- config 5: textured microfacet scene with cut-outs + mirror/glass/metal objects -------------------
	std::shared_ptr<Image2D> makeTexture(uint32_t n, int kind)
	{
		std::shared_ptr<Image2D> img = std::make_shared<Image2D>(n, n, 0xff000000u);
		for (uint32_t y = 0; y < n; ++y)
			for (uint32_t x = 0; x < n; ++x)
			{
				const float u = (float)x / (float)n, v = (float)y / (float)n;
				Pixel p(0.0f, 0.0f, 0.0f, 1.0f);
				if (kind == 0)        // albedo, sRGB-encoded checker with round alpha holes
				{
					const int cx = (int)(u * 16.0f), cy = (int)(v * 16.0f);
					const bool odd = ((cx + cy) & 1) != 0;
					p = odd ? Pixel(0.85f, 0.80f, 0.70f, 1.0f) : Pixel(0.55f, 0.25f, 0.20f, 1.0f);
					const float fu = u * 16.0f - (float)cx - 0.5f, fv = v * 16.0f - (float)cy - 0.5f;
					if (odd && fu * fu + fv * fv < 0.06f) p.a = 0.2f;
				}
				else if (kind == 1)   // tangent-space normal map: gentle ripples
				{
					const float nx = 0.25f * sinf(u * 100.5f), ny = 0.25f * cosf(v * 88.0f);
					const float nz = sqrtf(std::max(0.0f, 1.0f - nx * nx - ny * ny));
					p = Pixel(0.5f + 0.5f * nx, 0.5f + 0.5f * ny, 0.5f + 0.5f * nz, 1.0f);
				}
				else                  // roughness
				{
					const float r = 0.25f + 0.5f * (0.5f + 0.5f * sinf(u * 37.0f) * sinf(v * 41.0f));
					p = Pixel(r, r, r, 1.0f);
				}
				img->SetPixel((int32)x, (int32)y, p);
			}
		return img;
	}

	void buildTexturedRoom(SceneHandle scene, Owned& own, int G, int numBlobs, int icoLevel)
	{
		Gen g(11);
		own.images.push_back(makeTexture(1024, 0));
		own.images.push_back(makeTexture(1024, 1));
		own.images.push_back(makeTexture(1024, 2));
		MicrofacetMaterial* floorMat = own.mat(new MicrofacetMaterial);
		floorMat->SetAlbedoTexture(own.images[0]);
		floorMat->SetNormalTexture(own.images[1]);
		floorMat->SetRoughnessTexture(own.images[2]);
		MicrofacetMaterial* blobTextured = own.mat(new MicrofacetMaterial);
		blobTextured->SetAlbedoTexture(own.images[0]);
		blobTextured->SetRoughnessFallback(0.45f);
		MicrofacetMaterial* goldish = own.mat(MicrofacetMaterial::FromConstants(vec3(0.9f, 0.7f, 0.3f), 0.25f, 1.0f, vec3(0.0f)));
		MicrofacetMaterial* glow = own.mat(MicrofacetMaterial::FromConstants(vec3(0.2f, 0.2f, 0.2f), 0.8f, 0.0f, vec3(2.0f, 1.6f, 1.0f)));
		Material* mirror = own.mat(new Mirror(vec3(0.95f, 0.95f, 0.95f)));
		Material* glass = own.mat(new Dielectric(1.5f, vec3(0.95f, 1.0f, 0.95f)));
		Material* metal = own.mat(new Metal(vec3(0.8f, 0.8f, 0.9f), 0.15f));
		Material* wall = own.mat(new Lambertian(vec3(0.7f, 0.7f, 0.72f)));
		Material* variety[6] = { blobTextured, goldish, mirror, glass, metal, glow };

		StaticMesh* floor = own.keep(makeDisplacedGrid(G, 5.0f, 0.0f, floorMat, false, 8.0f));
		finalizeMesh(floor);
		Raylib_AddSceneElement(scene, (SceneElementHandle)floor);

		StaticMesh* room = own.keep(new StaticMesh);
		addQuad(room, vec3(-5, -1, -5), vec3(5, -1, -5), vec3(5, 6, -5), vec3(-5, 6, -5), vec3(0, 0, 1), wall);
		addQuad(room, vec3(-5, -1, 5), vec3(-5, -1, -5), vec3(-5, 6, -5), vec3(-5, 6, 5), vec3(1, 0, 0), wall);
		addQuad(room, vec3(5, -1, -5), vec3(5, -1, 5), vec3(5, 6, 5), vec3(5, 6, -5), vec3(-1, 0, 0), mirror);
		finalizeMesh(room);
		Raylib_AddSceneElement(scene, (SceneElementHandle)room);

		const IcoMesh ico = makeIcosphere(icoLevel);
		for (int k = 0; k < numBlobs; ++k)
		{
			const float s = g.range(0.25f, 0.6f);
			const vec3 center(g.range(-4.0f, 4.0f), 0.5f + s + g.range(0.0f, 1.5f), g.range(-4.0f, 3.0f));
			Material* m = variety[k % 6];
			StaticMesh* blob = own.keep(makeBlob(ico, center, vec3(s, s, s), 0.1f, g.range(0.0f, 6.28f), m, true));
			finalizeMesh(blob);
			Raylib_AddSceneElement(scene, (SceneElementHandle)blob);
		}
		Raylib_AddSceneElement(scene, (SceneElementHandle)own.keep(new Sphere(vec3(0.0f, 1.4f, 0.0f), 0.7f, glass)));
		Raylib_AddSceneElement(scene, (SceneElementHandle)own.keep(new Sphere(vec3(-2.0f, 1.2f, 1.5f), 0.5f, metal)));
		own.sky = makeGradientSky(512, 256, vec3(1.0f, 0.95f, 0.9f), vec3(0.4f, 0.6f, 1.0f));
		Raylib_SetSkyPanorama(scene, own.sky);
		Raylib_SetSunIlluminance(scene, 12.0f, 11.0f, 10.0f);
		Raylib_SetSunDirection(scene, -0.3f, -1.0f, -0.4f);
	}
}

// ---- config 8: OBJ + MTL through the object API (reference conversion rules) ----------------------------------
namespace
{
	struct ObjMtl { std::string name; float Kd[3] = { 0, 0, 0 }, Ks[3] = { 0, 0, 0 }, Ke[3] = { 0, 0, 0 }, Tf[3] = { 0, 0, 0 }; float Ns = 1.0f, Ni = 1.0f, Pr = 0.0f, Pm = 0.0f; int illum = 0; };
	struct ObjCorner { int v, vt, vn; };
	struct ObjShape { std::vector<ObjCorner> corners; std::vector<int> faceMaterial; };

	int objIndex(int i, size_t count) { return i > 0 ? i - 1 : (i < 0 ? (int)count + i : -1); }

	bool buildFromObj(SceneHandle scene, Owned& own, const char* path)
	{
		FILE* f = fopen(path, "r");
		if (!f) return false;
		std::string dir(path);
		const size_t slash = dir.find_last_of('/');
		dir = slash == std::string::npos ? std::string() : dir.substr(0, slash + 1);
		std::vector<float> P, T, N;
		std::vector<ObjMtl> mtls;
		std::vector<ObjShape> shapes;
		ObjShape cur;
		int curMtl = -1;
		char line[1024];
		while (fgets(line, sizeof(line), f))
		{
			char key[64] = { 0 };
			if (sscanf(line, "%63s", key) != 1 || key[0] == '#') continue;
			const char* rest = strstr(line, key) + strlen(key);
			if (!strcmp(key, "v")) { float a = 0, b = 0, c = 0; sscanf(rest, "%f %f %f", &a, &b, &c); P.push_back(a); P.push_back(b); P.push_back(c); }
			else if (!strcmp(key, "vt")) { float a = 0, b = 0; sscanf(rest, "%f %f", &a, &b); T.push_back(a); T.push_back(b); }
			else if (!strcmp(key, "vn")) { float a = 0, b = 0, c = 0; sscanf(rest, "%f %f %f", &a, &b, &c); N.push_back(a); N.push_back(b); N.push_back(c); }
			else if (!strcmp(key, "g") || !strcmp(key, "o")) { if (!cur.faceMaterial.empty()) shapes.push_back(cur); cur = ObjShape(); }
			else if (!strcmp(key, "usemtl"))
			{
				char name[256] = { 0 };
				sscanf(rest, "%255s", name);
				curMtl = -1;
				for (size_t i = 0; i < mtls.size(); ++i) if (mtls[i].name == name) curMtl = (int)i;
			}
			else if (!strcmp(key, "mtllib"))
			{
				char name[256] = { 0 };
				sscanf(rest, "%255s", name);
				FILE* m = fopen((dir + name).c_str(), "r");
				if (!m) continue;
				char ml[1024];
				while (fgets(ml, sizeof(ml), m))
				{
					char mk[64] = { 0 };
					if (sscanf(ml, "%63s", mk) != 1 || mk[0] == '#') continue;
					const char* mr = strstr(ml, mk) + strlen(mk);
					if (!strcmp(mk, "newmtl")) { mtls.push_back(ObjMtl()); char nm[256] = { 0 }; sscanf(mr, "%255s", nm); mtls.back().name = nm; continue; }
					if (mtls.empty()) continue;
					ObjMtl& M = mtls.back();
					if (!strcmp(mk, "Kd")) sscanf(mr, "%f %f %f", &M.Kd[0], &M.Kd[1], &M.Kd[2]);
					else if (!strcmp(mk, "Ks")) sscanf(mr, "%f %f %f", &M.Ks[0], &M.Ks[1], &M.Ks[2]);
					else if (!strcmp(mk, "Ke")) sscanf(mr, "%f %f %f", &M.Ke[0], &M.Ke[1], &M.Ke[2]);
					else if (!strcmp(mk, "Tf") || !strcmp(mk, "Kt")) sscanf(mr, "%f %f %f", &M.Tf[0], &M.Tf[1], &M.Tf[2]);
					else if (!strcmp(mk, "Ns")) sscanf(mr, "%f", &M.Ns);
					else if (!strcmp(mk, "Ni")) sscanf(mr, "%f", &M.Ni);
					else if (!strcmp(mk, "Pr")) sscanf(mr, "%f", &M.Pr);
					else if (!strcmp(mk, "Pm")) sscanf(mr, "%f", &M.Pm);
					else if (!strcmp(mk, "illum")) sscanf(mr, "%d", &M.illum);
				}
				fclose(m);
			}
			else if (!strcmp(key, "f"))
			{
				ObjCorner c[3];
				int n = 0;
				const char* p = rest;
				while (n < 3)
				{
					while (*p == ' ' || *p == '\t') ++p;
					if (!*p || *p == '\n' || *p == '\r') break;
					int v = 0, vt = 0, vn = 0;
					if (sscanf(p, "%d/%d/%d", &v, &vt, &vn) == 3) {}
					else if (sscanf(p, "%d//%d", &v, &vn) == 2) { vt = 0; }
					else if (sscanf(p, "%d/%d", &v, &vt) == 2) { vn = 0; }
					else { sscanf(p, "%d", &v); vt = vn = 0; }
					c[n].v = objIndex(v, P.size() / 3); c[n].vt = vt ? objIndex(vt, T.size() / 2) : -1; c[n].vn = vn ? objIndex(vn, N.size() / 3) : -1;
					++n;
					while (*p && *p != ' ' && *p != '\t' && *p != '\n') ++p;
				}
				if (n == 3) { cur.corners.push_back(c[0]); cur.corners.push_back(c[1]); cur.corners.push_back(c[2]); cur.faceMaterial.push_back(curMtl); }
			}
		}
		fclose(f);
		if (!cur.faceMaterial.empty()) shapes.push_back(cur);
		if (shapes.empty()) return false;

		// materials (obj_loader.cc:342-398)
		std::vector<Material*> materials;
		for (const ObjMtl& M : mtls)
		{
			const vec3 albedo = min(vec3(0.95f), vec3(M.Kd[0], M.Kd[1], M.Kd[2]));
			const bool transparent = M.illum == 4 || M.illum == 6;
			if (transparent && albedo == vec3(0.0f)) materials.push_back(own.mat(new Dielectric(M.Ni, vec3(M.Tf[0], M.Tf[1], M.Tf[2]))));
			else if (M.illum == 3) materials.push_back(own.mat(new Mirror(albedo)));
			else
			{
				MicrofacetMaterial* mf = new MicrofacetMaterial;
				mf->SetAlbedoFallback(albedo);
				if (M.Pr > 0.0f) mf->SetRoughnessFallback(M.Pr);
				else
				{
					const float intensity = (M.Ks[0] + M.Ks[1] + M.Ks[2]) / 3.0f;
					mf->SetRoughnessFallback(std::sqrt(2.0f / (M.Ns * intensity + 2.0f)));
				}
				mf->SetMetallicFallback(M.Pm);
				mf->SetEmissiveFallback(vec3(M.Ke[0], M.Ke[1], M.Ke[2]));
				materials.push_back(own.mat(mf));
			}
		}
		Material* fallback = own.mat(new Lambertian(vec3(0.5f, 0.5f, 0.5f)));

		// shapes -> meshes (obj_loader.cc:133-228), root (:230-241)
		std::vector<Hitable*> roots;
		for (const ObjShape& shape : shapes)
		{
			StaticMesh* mesh = own.keep(new StaticMesh);
			for (size_t face = 0; face < shape.faceMaterial.size(); ++face)
			{
				vec3 p[3], n[3];
				float us[3], vs[3];
				bool validNormal = true;
				for (int k = 0; k < 3; ++k)
				{
					const ObjCorner& c = shape.corners[3 * face + k];
					p[k] = vec3(P[3 * c.v], P[3 * c.v + 1], P[3 * c.v + 2]);
					us[k] = vs[k] = 0.0f;
					if (c.vt >= 0) { us[k] = T[2 * c.vt]; vs[k] = T[2 * c.vt + 1]; }
					n[k] = vec3(0.0f, 0.0f, 0.0f);
					if (c.vn >= 0) n[k] = vec3(N[3 * c.vn], N[3 * c.vn + 1], N[3 * c.vn + 2]);
					else validNormal = false;
				}
				if (!validNormal) { const vec3 fn = cross(p[1] - p[0], p[2] - p[0]); n[0] = n[1] = n[2] = normalize(fn); }
				const int mid = shape.faceMaterial[face];
				Material* faceMaterial = (mid >= 0 && mid < (int)materials.size()) ? materials[mid] : fallback;
				Triangle tri(p[0], p[1], p[2], n[0], n[1], n[2], faceMaterial);
				tri.SetParameterization(us[0], vs[0], us[1], vs[1], us[2], vs[2]);
				countedAdd(mesh, tri);
			}
			mesh->CalculateBounds();
			finalizeMesh(mesh);
			roots.push_back(mesh);
		}
		Hitable* root = roots[0];
		if (roots.size() > 1)
		{
			HitableList* list = own.keep(new HitableList(roots));
			scene_hook_before_bvh_build();
			root = own.keep(new BVHNode(list, 0.0f, 0.0f));
		}
		Raylib_AddSceneElement(scene, (SceneElementHandle)root);
		return true;
	}
}

extern "C" {

struct DemoSceneInfo
{
	SceneHandle  scene;
	CameraHandle camera;
	RendererSettings settings;     // the BASELINE.json configuration (viewport, spp, depth, tMin, mode)
	uint64_t numTriangles, numSpheres, numMeshes;
	float cameraPos[3], cameraLookAt[3];
	float fovY, aperture, focalDistance, shutterBegin, shutterEnd;
};

// config: 1..5 = BASELINE.json configs[0..4]; 6 = tiny mixed scene (spheres + cube + triangles + all materials);
// 7 = a raw HitableList scene element (spheres, cubes, triangles, duplicated members) next to ordinary elements;
// 8 = the OBJ + MTL file named by $DEMO_OBJ_PATH, imported through the object API with the reference's conversion rules.
// sizeParam: 0 = the configuration's own size; otherwise grid resolution G (3, 5), instance count (4).
__attribute__((visibility("default")))
int32_t demo_scene_create(int32_t config, int32_t sizeParam, DemoSceneInfo* out)
{
	if (!out) return 0;
	memset(out, 0, sizeof(*out));
	g_meshTriangles = 0;
	SceneHandle scene = Raylib_CreateScene();
	Owned* own = new Owned;
	vec3 camPos(0.0f), camAt(0.0f, 0.0f, -1.0f);
	float fov = 60.0f, aperture = 0.0f, t0 = 0.0f, t1 = 0.0f;
	RendererSettings rs;
	rs.viewportWidth = 640; rs.viewportHeight = 360; rs.samplesPerPixel = 16; rs.maxPathLength = 5; rs.rayTMin = 0.0001f;
	rs.renderMode = RAYLIB_RENDERMODE_Default;
	switch (config)
	{
	case 1:
		buildRandomSpheres(scene, *own);
		camPos = vec3(0.0f, 1.5f, 5.0f); camAt = vec3(0.0f, 0.5f, 0.0f); fov = 60.0f; aperture = 0.01f; t0 = 0.0f; t1 = 5.0f;
		break;
	case 2:
		buildCornell(scene, *own);
		camPos = vec3(0.0f, 1.0f, 4.0f); camAt = vec3(0.0f, 1.0f, -1.0f); fov = 45.0f;
		rs.viewportWidth = 1920; rs.viewportHeight = 1080; rs.samplesPerPixel = 64; rs.maxPathLength = 8;
		break;
	case 3:
		buildDisplacedGrid(scene, *own, sizeParam > 0 ? sizeParam : 708);
		camPos = vec3(0.0f, 4.0f, 8.0f); camAt = vec3(0.0f, 0.0f, 0.0f); fov = 60.0f;
		rs.viewportWidth = 1920; rs.viewportHeight = 1080; rs.samplesPerPixel = 16; rs.maxPathLength = 2;
		break;
	case 4:
		buildInstanceScatter(scene, *own, sizeParam > 0 ? sizeParam : 7812, 3);
		{
			int n = sizeParam > 0 ? sizeParam : 7812, side = 1; while (side * side < n) ++side;
			const float half = 0.5f * (float)side;
			camPos = vec3(0.0f, 0.35f * half + 2.0f, 1.25f * half + 3.0f); camAt = vec3(0.0f, 0.0f, 0.0f); fov = 45.0f;
		}
		rs.viewportWidth = 3840; rs.viewportHeight = 2160; rs.samplesPerPixel = 256; rs.maxPathLength = 8;
		break;
	case 5:
		buildTexturedRoom(scene, *own, sizeParam > 0 ? sizeParam : 900, sizeParam > 0 ? std::max(6, sizeParam / 12) : 75, sizeParam > 0 && sizeParam < 200 ? 2 : 4);
		camPos = vec3(0.5f, 2.5f, 4.6f); camAt = vec3(0.0f, 1.0f, 0.0f); fov = 60.0f; aperture = 0.0f; t0 = 0.0f; t1 = 1.0f;
		rs.viewportWidth = 1920; rs.viewportHeight = 1080; rs.samplesPerPixel = 64; rs.maxPathLength = 8;
		break;
	case 6:
	{
		Material* lam = own->mat(new Lambertian(vec3(0.6f, 0.3f, 0.2f)));
		Material* met = own->mat(new Metal(vec3(0.8f, 0.8f, 0.7f), 0.2f));
		Material* gls = own->mat(new Dielectric(1.4f));
		Material* mir = own->mat(new Mirror(vec3(0.9f, 0.9f, 0.9f)));
		Material* lit = own->mat(new DiffuseLight(vec3(4.0f, 3.5f, 3.0f)));
		Material* mic = own->mat(MicrofacetMaterial::FromConstants(vec3(0.7f, 0.7f, 0.2f), 0.4f, 0.3f, vec3(0.05f, 0.0f, 0.0f)));
		auto addS = [&](const vec3& c, float r, Material* m) { Raylib_AddSceneElement(scene, (SceneElementHandle)own->keep(new Sphere(c, r, m))); };
		addS(vec3(0.0f, -100.5f, -1.0f), 100.0f, mic);
		addS(vec3(-1.1f, 0.0f, -1.0f), 0.5f, gls);
		addS(vec3(0.0f, 0.0f, -1.0f), 0.5f, lam);
		addS(vec3(1.1f, 0.0f, -1.0f), 0.5f, met);
		addS(vec3(0.0f, 1.6f, -1.0f), 0.4f, lit);
		Raylib_AddSceneElement(scene, (SceneElementHandle)own->keep(new Cube(vec3(-2.4f, -0.5f, -2.0f), vec3(-1.8f, 0.3f, -1.4f), 0.0f, vec3(0.0f, 0.1f, 0.0f), lam)));
		const vec3 n(0.0f, 0.0f, 1.0f);
		Raylib_AddSceneElement(scene, (SceneElementHandle)own->keep(new Triangle(vec3(-2.5f, -0.5f, -2.5f), vec3(2.5f, -0.5f, -2.5f), vec3(2.5f, 2.0f, -2.5f), n, n, n, mir)));
		Raylib_AddSceneElement(scene, (SceneElementHandle)own->keep(new Triangle(vec3(-2.5f, -0.5f, -2.5f), vec3(2.5f, 2.0f, -2.5f), vec3(-2.5f, 2.0f, -2.5f), n, n, n, mir)));
		StaticMesh* mesh = own->keep(new StaticMesh);
		addBox(mesh, vec3(1.9f, -0.2f, -0.3f), vec3(0.25f, 0.3f, 0.25f), 30.0f, met);
		finalizeMesh(mesh);
		Raylib_AddSceneElement(scene, (SceneElementHandle)mesh);
		own->sky = makeGradientSky(64, 32, vec3(1.0f, 1.0f, 1.0f), vec3(0.5f, 0.7f, 1.0f));
		Raylib_SetSkyPanorama(scene, own->sky);
		Raylib_SetSunIlluminance(scene, 3.0f, 3.0f, 2.5f);
		Raylib_SetSunDirection(scene, -0.4f, -1.0f, -0.3f);
		camPos = vec3(0.0f, 0.6f, 3.2f); camAt = vec3(0.0f, 0.2f, -1.0f); fov = 50.0f; aperture = 0.02f; t0 = 0.0f; t1 = 2.0f;
		rs.viewportWidth = 320; rs.viewportHeight = 180; rs.samplesPerPixel = 8; rs.maxPathLength = 6;
		break;
	}
	case 7:
	{
		// A raw HitableList as a scene element (geom/hit.cc:34-50): spheres, cubes and triangles scanned linearly with a
		// shrinking upper bound, duplicates included so that equal-t ties between members occur on many rays.  The list's
		// BoundingBox never writes its out-parameter in the reference (geom/hit.h:62-86), so the BVH build sees it as a point
		// at the origin: the members sit around the origin, where the box of the node holding the list is.
		Material* lam = own->mat(new Lambertian(vec3(0.7f, 0.4f, 0.3f)));
		Material* met = own->mat(new Metal(vec3(0.8f, 0.8f, 0.9f), 0.1f));
		Material* lit = own->mat(new DiffuseLight(vec3(3.0f, 3.0f, 3.0f)));
		Material* mir = own->mat(new Mirror(vec3(0.9f, 0.8f, 0.8f)));
		const vec3 n(0.0f, 0.0f, 1.0f);
		std::vector<Hitable*> members;
		members.push_back(own->keep(new Triangle(vec3(-1.2f, -0.6f, -0.4f), vec3(1.2f, -0.6f, -0.4f), vec3(1.2f, 0.9f, -0.4f), n, n, n, lam)));
		members.push_back(own->keep(new Sphere(vec3(-0.45f, 0.1f, 0.1f), 0.3f, met)));
		members.push_back(own->keep(new Sphere(vec3(-0.45f, 0.1f, 0.1f), 0.3f, lit)));              // same sphere again: the first wins
		members.push_back(own->keep(new Triangle(vec3(-1.2f, -0.6f, -0.4f), vec3(1.2f, -0.6f, -0.4f), vec3(1.2f, 0.9f, -0.4f), n, n, n, mir)));   // same triangle again: the later wins
		members.push_back(own->keep(new Cube(vec3(0.2f, -0.5f, -0.2f), vec3(0.7f, 0.0f, 0.3f), 0.0f, vec3(0.0f, 0.0f, 0.0f), lam)));
		members.push_back(own->keep(new Triangle(vec3(-1.2f, -0.6f, -0.4f), vec3(1.2f, 0.9f, -0.4f), vec3(-1.2f, 0.9f, -0.4f), n, n, n, met)));
		members.push_back(own->keep(new Sphere(vec3(0.45f, 0.45f, 0.0f), 0.22f, lam)));
		members.push_back(own->keep(new Cube(vec3(0.2f, -0.5f, -0.2f), vec3(0.7f, 0.0f, 0.3f), 0.0f, vec3(0.0f, 0.0f, 0.0f), met)));     // same cube again: the later wins
		HitableList* list = own->keep(new HitableList(members));
		Raylib_AddSceneElement(scene, (SceneElementHandle)list);
		auto addS = [&](const vec3& c, float r, Material* m) { Raylib_AddSceneElement(scene, (SceneElementHandle)own->keep(new Sphere(c, r, m))); };
		addS(vec3(0.0f, -100.6f, 0.0f), 100.0f, lam);
		addS(vec3(-1.4f, -0.3f, 0.6f), 0.3f, met);
		addS(vec3(1.5f, -0.2f, 0.4f), 0.4f, mir);
		addS(vec3(0.0f, 1.5f, -0.2f), 0.35f, lit);
		own->sky = makeGradientSky(64, 32, vec3(1.0f, 1.0f, 1.0f), vec3(0.5f, 0.7f, 1.0f));
		Raylib_SetSkyPanorama(scene, own->sky);
		Raylib_SetSunIlluminance(scene, 2.0f, 2.0f, 2.0f);
		Raylib_SetSunDirection(scene, -0.3f, -1.0f, -0.5f);
		camPos = vec3(0.0f, 0.5f, 3.0f); camAt = vec3(0.0f, 0.1f, 0.0f); fov = 45.0f; aperture = 0.0f; t0 = 0.0f; t1 = 0.0f;
		rs.viewportWidth = 320; rs.viewportHeight = 180; rs.samplesPerPixel = 8; rs.maxPathLength = 5;
		break;
	}
	case 8:
	{
		// A Wavefront OBJ + MTL imported THROUGH THE CLIENT OBJECT API, following the reference's conversion rules
		// (loader/obj_loader.cc:113-245 shapes -> StaticMesh, :294-400 materials): the oracle for Raylib_LoadOBJModel, whose
		// reference implementation needs tinyobjloader and cannot be compiled here.  File: $DEMO_OBJ_PATH; the parser below
		// understands what the test writes (v / vt / vn / f with v, v/vt, v//vn or v/vt/vn triangles, g, usemtl, mtllib;
		// newmtl, Kd, Ks, Ke, Tf, Ns, Ni, illum, Pr, Pm) and shares no code with the product's importer.
		const char* path = getenv("DEMO_OBJ_PATH");
		if (!path || !buildFromObj(scene, *own, path)) { delete own; Raylib_DestroyScene(scene); return 0; }
		Raylib_SetSunIlluminance(scene, 5.0f, 5.0f, 5.0f);
		Raylib_SetSunDirection(scene, 0.2f, -1.0f, -0.3f);
		camPos = vec3(0.3f, 1.6f, 3.4f); camAt = vec3(0.0f, 0.3f, 0.0f); fov = 55.0f; aperture = 0.0f; t0 = 0.0f; t1 = 0.0f;
		rs.viewportWidth = 240; rs.viewportHeight = 136; rs.samplesPerPixel = 4; rs.maxPathLength = 5;
		break;
	}
	default:
		delete own;
		Raylib_DestroyScene(scene);
		return 0;
	}
	finalizeScene(scene);

	const float focal = (camPos - camAt).Length();
	CameraHandle camera = Raylib_CreateCamera();
	Raylib_CameraSetPosition(camera, camPos.x, camPos.y, camPos.z);
	Raylib_CameraSetLookAt(camera, camAt.x, camAt.y, camAt.z);
	Raylib_CameraSetPerspective(camera, fov, rs.getViewportAspectWH());
	Raylib_CameraSetLens(camera, aperture, focal);
	Raylib_CameraSetMotion(camera, t0, t1);

	uint64_t tris = 0, spheres = 0, meshes = 0;
	for (Hitable* h : own->hitables)
	{
		if (dynamic_cast<Sphere*>(h)) spheres++;
		else if (dynamic_cast<Triangle*>(h)) tris++;
		else if (dynamic_cast<StaticMesh*>(h)) meshes++;
	}
	out->scene = scene; out->camera = camera; out->settings = rs;
	out->numTriangles = tris + g_meshTriangles; out->numSpheres = spheres; out->numMeshes = meshes;
	out->cameraPos[0] = camPos.x; out->cameraPos[1] = camPos.y; out->cameraPos[2] = camPos.z;
	out->cameraLookAt[0] = camAt.x; out->cameraLookAt[1] = camAt.y; out->cameraLookAt[2] = camAt.z;
	out->fovY = fov; out->aperture = aperture; out->focalDistance = focal; out->shutterBegin = t0; out->shutterEnd = t1;

	std::lock_guard<std::mutex> lock(g_mutex);
	g_owned[scene] = own;
	return 1;
}

__attribute__((visibility("default")))
void demo_scene_destroy(SceneHandle scene, CameraHandle camera)
{
	Owned* own = nullptr;
	{
		std::lock_guard<std::mutex> lock(g_mutex);
		auto it = g_owned.find(scene);
		if (it != g_owned.end()) { own = it->second; g_owned.erase(it); }
	}
	scene_hook_on_destroy(scene);
	Raylib_DestroyScene(scene);     // drops the (GPU copy and) top-level BVH before the elements go away
	if (camera) Raylib_DestroyCamera(camera);
	delete own;
}

// Aspect must follow the viewport when a test renders a configuration at a reduced resolution.
__attribute__((visibility("default")))
void demo_camera_set_aspect(CameraHandle camera, float fovY, uint32_t width, uint32_t height)
{
	Raylib_CameraSetPerspective(camera, fovY, (float)width / (float)height);
}

} // extern "C"
