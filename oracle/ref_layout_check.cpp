// Compile-time comparison of the device-visible structs this repository defines with the REFERENCE'S OWN definitions,
// for the reference headers that compile standalone (PrimitiveSet.h, BoundingSphere.h with the vendored GLM; everything that
// includes vulkan.hpp does not - DESIGN.md section 6).  TEST INFRASTRUCTURE, built by `make -C oracle ref` in the build
// container only: if this file compiles, the layouts agree; the binary just says so.
#include <CadR/PrimitiveSet.h>
#include <CadR/BoundingSphere.h>
#include "../include/cadr_b200.h"
#include <cstddef>
#include <cstdio>

// PrimitiveSet (src/CadR/PrimitiveSet.h:12-15) == cadr_primitive_set: what lookupHandle(primitiveSetHandle) + offset points at
static_assert(sizeof(CadR::PrimitiveSet) == sizeof(cadr_primitive_set) && sizeof(cadr_primitive_set) == 8, "PrimitiveSet size");
static_assert(offsetof(CadR::PrimitiveSet, indexCount) == offsetof(cadr_primitive_set, count), "PrimitiveSet::indexCount");
static_assert(offsetof(CadR::PrimitiveSet, startIndex) == offsetof(cadr_primitive_set, first), "PrimitiveSet::startIndex");

// BoundingSphere (src/CadR/BoundingSphere.h:18-21) == cadr_drawable_cull_data::sphere {x, y, z, radius}
static_assert(sizeof(CadR::BoundingSphere) == sizeof(cadr_drawable_cull_data::sphere) && sizeof(CadR::BoundingSphere) == 16, "BoundingSphere size");
static_assert(offsetof(CadR::BoundingSphere, center) == 0 && offsetof(CadR::BoundingSphere, radius) == 12, "BoundingSphere members");
static_assert(offsetof(cadr_drawable_cull_data, sphere) == 0, "sphere leads the culling record");

// glm::mat4 is 64 bytes, column-major: the MatrixList payload (MatrixList.h:54-59)
static_assert(sizeof(glm::mat4) == CADR_MATRIX_BYTES, "mat4 size");

int main()
{
	glm::mat4 m(1.f);
	m[3] = glm::vec4(7.f, 8.f, 9.f, 1.f);      // column 3 = translation
	const float* f = &m[0][0];
	if(f[12] != 7.f || f[13] != 8.f || f[14] != 9.f) { fprintf(stderr, "glm::mat4 is not column-major\n"); return 1; }
	// empty sphere convention (BoundingSphere.h:39-43): radius -inf, i.e. < 0 as the culling record documents
	if(!(CadR::BoundingSphere::empty().radius < 0.f) || !CadR::BoundingSphere::empty().isEmpty()) return 1;
	printf("layout_check ok\n");
	return 0;
}
