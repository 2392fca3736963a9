// Runs the REFERENCE'S OWN bounding-sphere transform - `operator*(const glm::mat4&, BoundingSphere)` of
// /root/reference/src/CadR/BoundingSphere.h:70-87, compiled from where it lies together with the reference's vendored
// GLM - on (matrix, sphere) pairs.  TEST INFRASTRUCTURE: oracle/make_golden.py uses it to produce
// tests/golden/bounding_sphere_ref.npz, the only step of the culling extension (Tier X) that has a counterpart in
// the reference.
//
//   stdin : n records of 80 bytes: column-major mat4 (16 floats) + sphere {cx, cy, cz, r}
//   stdout: n records of 16 bytes: transformed sphere {cx, cy, cz, r}
#include <CadR/BoundingSphere.h>
#include <cstdio>
#include <cstring>

int main()
{
	float rec[20];
	while(fread(rec, sizeof(float), 20, stdin) == 20) {
		glm::mat4 m;
		memcpy(&m, rec, 64);
		CadR::BoundingSphere bs{glm::vec3(rec[16], rec[17], rec[18]), rec[19]};
		CadR::BoundingSphere w = m * bs;
		float out[4] = {w.center.x, w.center.y, w.center.z, w.radius};
		fwrite(out, sizeof(float), 4, stdout);
	}
	return 0;
}
