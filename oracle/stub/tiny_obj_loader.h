// oracle/stub/tiny_obj_loader.h -- TEST INFRASTRUCTURE. tinyobjloader
// v2.0.0rc10 (Setup.ps1:39) is not vendored; raylib/loader/obj_loader.h:49-58
// only needs these two names to DECLARE OBJLoader.  OBJ parsing is out of scope.
#pragma once
namespace tinyobj { struct material_t {}; struct ObjReader {}; }
