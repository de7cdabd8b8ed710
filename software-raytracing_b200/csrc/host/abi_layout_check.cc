// abi_layout_check.cc -- pins the object sizes of the reference's public classes (measured with
// g++ on the reference headers, SURVEY.md section 7 "ABI breadth").  A client binary compiled
// against either header set must agree with this library on every one of them.
#include "geom/scene.h"
#include "geom/primitives.h"
#include "render/camera.h"
#include "render/material.h"
#include "render/image.h"
#include "raylib_types.h"
#include <cstddef>

static_assert(sizeof(vec3) == 12, "vec3");
static_assert(sizeof(ray) == 28, "ray");
static_assert(sizeof(AABB) == 24, "AABB");
static_assert(sizeof(HitResult) == 72, "HitResult");
static_assert(sizeof(BVHNode) == 48, "BVHNode");
static_assert(sizeof(Sphere) == 32, "Sphere");
static_assert(sizeof(Cube) == 56, "Cube");
static_assert(sizeof(Triangle) == 152, "Triangle");
static_assert(sizeof(StaticMesh) == 80, "StaticMesh");
static_assert(sizeof(Scene) == 80, "Scene");
static_assert(sizeof(Camera) == 128, "Camera");
static_assert(sizeof(Lambertian) == 24, "Lambertian");
static_assert(sizeof(Metal) == 24, "Metal");
static_assert(sizeof(Dielectric) == 24, "Dielectric");
static_assert(sizeof(Mirror) == 24, "Mirror");
static_assert(sizeof(DiffuseLight) == 24, "DiffuseLight");
static_assert(sizeof(MicrofacetMaterial) == 80, "MicrofacetMaterial");
static_assert(sizeof(RendererSettings) == 24, "RendererSettings");
static_assert(sizeof(Pixel) == 16, "Pixel");
static_assert(sizeof(Image2D) == 32, "Image2D");

// Field offsets, measured with offsetof on the REFERENCE's own headers (g++ 13, x86-64; the probe printed these very
// lines): a client that fills or reads these objects in place -- src/main.cc:597-983 news them with inline constructors --
// must find every member where this library looks for it.  This file is compiled with -fno-access-control.
static_assert(offsetof(HitResult, t) == 0, "HitResult::t");
static_assert(offsetof(HitResult, p) == 4, "HitResult::p");
static_assert(offsetof(HitResult, n) == 16, "HitResult::n");
static_assert(offsetof(HitResult, paramU) == 28, "HitResult::paramU");
static_assert(offsetof(HitResult, paramV) == 32, "HitResult::paramV");
static_assert(offsetof(HitResult, material) == 40, "HitResult::material");
static_assert(offsetof(HitResult, tangent) == 48, "HitResult::tangent");
static_assert(offsetof(HitResult, bitangent) == 60, "HitResult::bitangent");
static_assert(offsetof(BVHNode, left) == 8, "BVHNode::left");
static_assert(offsetof(BVHNode, right) == 16, "BVHNode::right");
static_assert(offsetof(BVHNode, box) == 24, "BVHNode::box");
static_assert(offsetof(Sphere, center) == 8, "Sphere::center");
static_assert(offsetof(Sphere, radius) == 20, "Sphere::radius");
static_assert(offsetof(Sphere, material) == 24, "Sphere::material");
static_assert(offsetof(Cube, minBounds) == 8, "Cube::minBounds");
static_assert(offsetof(Cube, maxBounds) == 20, "Cube::maxBounds");
static_assert(offsetof(Cube, timeStartMove) == 32, "Cube::timeStartMove");
static_assert(offsetof(Cube, velocity) == 36, "Cube::velocity");
static_assert(offsetof(Cube, material) == 48, "Cube::material");
static_assert(offsetof(Triangle, v0) == 8, "Triangle::v0");
static_assert(offsetof(Triangle, v1) == 20, "Triangle::v1");
static_assert(offsetof(Triangle, v2) == 32, "Triangle::v2");
static_assert(offsetof(Triangle, n) == 44, "Triangle::n");
static_assert(offsetof(Triangle, n0) == 56, "Triangle::n0");
static_assert(offsetof(Triangle, n1) == 68, "Triangle::n1");
static_assert(offsetof(Triangle, n2) == 80, "Triangle::n2");
static_assert(offsetof(Triangle, s0) == 116, "Triangle::s0");
static_assert(offsetof(Triangle, t0) == 120, "Triangle::t0");
static_assert(offsetof(Triangle, s1) == 124, "Triangle::s1");
static_assert(offsetof(Triangle, t1) == 128, "Triangle::t1");
static_assert(offsetof(Triangle, s2) == 132, "Triangle::s2");
static_assert(offsetof(Triangle, t2) == 136, "Triangle::t2");
static_assert(offsetof(Triangle, material) == 144, "Triangle::material");
static_assert(offsetof(StaticMesh, triangles) == 8, "StaticMesh::triangles");
static_assert(offsetof(StaticMesh, bvh) == 64, "StaticMesh::bvh");
static_assert(offsetof(StaticMesh, bounds) == 32, "StaticMesh::bounds");
static_assert(offsetof(StaticMesh, bLocked) == 72, "StaticMesh::bLocked");
static_assert(offsetof(Scene, hitableList) == 0, "Scene::hitableList");
static_assert(offsetof(Scene, accelStruct) == 32, "Scene::accelStruct");
static_assert(offsetof(Scene, skyPanorama) == 40, "Scene::skyPanorama");
static_assert(offsetof(Scene, sunIlluminance) == 48, "Scene::sunIlluminance");
static_assert(offsetof(Scene, sunDirection) == 60, "Scene::sunDirection");
static_assert(offsetof(Scene, bFinalized) == 72, "Scene::bFinalized");
static_assert(offsetof(Camera, origin) == 0, "Camera::origin");
static_assert(offsetof(Camera, lookAt) == 12, "Camera::lookAt");
static_assert(offsetof(Camera, fovY_degrees) == 24, "Camera::fovY_degrees");
static_assert(offsetof(Camera, aspectWH) == 28, "Camera::aspectWH");
static_assert(offsetof(Camera, aperture) == 32, "Camera::aperture");
static_assert(offsetof(Camera, focalDistance) == 36, "Camera::focalDistance");
static_assert(offsetof(Camera, beginTime) == 40, "Camera::beginTime");
static_assert(offsetof(Camera, endTime) == 44, "Camera::endTime");
static_assert(offsetof(Lambertian, albedo) == 8, "Lambertian::albedo");
static_assert(offsetof(Metal, albedo) == 8, "Metal::albedo");
static_assert(offsetof(Metal, fuzziness) == 20, "Metal::fuzziness");
static_assert(offsetof(Dielectric, ref_idx) == 8, "Dielectric::ref_idx");
static_assert(offsetof(Dielectric, transmissionFilter) == 12, "Dielectric::transmissionFilter");
static_assert(offsetof(Mirror, baseColor) == 8, "Mirror::baseColor");
static_assert(offsetof(DiffuseLight, intensity) == 8, "DiffuseLight::intensity");
static_assert(offsetof(MicrofacetMaterial, albedoTexture) == 8, "MicrofacetMaterial::albedoTexture");
static_assert(offsetof(MicrofacetMaterial, normalmapTexture) == 16, "MicrofacetMaterial::normalmapTexture");
static_assert(offsetof(MicrofacetMaterial, roughnessTexture) == 24, "MicrofacetMaterial::roughnessTexture");
static_assert(offsetof(MicrofacetMaterial, metallicTexture) == 32, "MicrofacetMaterial::metallicTexture");
static_assert(offsetof(MicrofacetMaterial, emissiveTexture) == 40, "MicrofacetMaterial::emissiveTexture");
static_assert(offsetof(MicrofacetMaterial, albedoFallback) == 48, "MicrofacetMaterial::albedoFallback");
static_assert(offsetof(MicrofacetMaterial, roughnessFallback) == 60, "MicrofacetMaterial::roughnessFallback");
static_assert(offsetof(MicrofacetMaterial, metallicFallback) == 64, "MicrofacetMaterial::metallicFallback");
static_assert(offsetof(MicrofacetMaterial, emissiveFallback) == 68, "MicrofacetMaterial::emissiveFallback");
static_assert(offsetof(Image2D, width) == 0, "Image2D::width");
static_assert(offsetof(Image2D, height) == 4, "Image2D::height");
static_assert(offsetof(Image2D, image) == 8, "Image2D::image");
static_assert(offsetof(RendererSettings, viewportWidth) == 0, "RendererSettings::viewportWidth");
static_assert(offsetof(RendererSettings, viewportHeight) == 4, "RendererSettings::viewportHeight");
static_assert(offsetof(RendererSettings, samplesPerPixel) == 8, "RendererSettings::samplesPerPixel");
static_assert(offsetof(RendererSettings, maxPathLength) == 12, "RendererSettings::maxPathLength");
static_assert(offsetof(RendererSettings, rayTMin) == 16, "RendererSettings::rayTMin");
static_assert(offsetof(RendererSettings, renderMode) == 20, "RendererSettings::renderMode");
