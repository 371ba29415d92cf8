// TEST INFRASTRUCTURE: glm's own mix / slerp (the calls the reference's animation block makes, niagara.cpp:1386-1388), exported
// for property tests of nvc_host_animate.  Built with the reference's glm and its defines into oracle/_ref/libanim_ref.so.
#include "host_golden.h"

extern "C" void anim_ref_blend(const float* t0, float s0, const float* r0_xyzw, const float* t1, float s1, const float* r1_xyzw, float a, float* out8)
{
	vec3 p = glm::mix(vec3(t0[0], t0[1], t0[2]), vec3(t1[0], t1[1], t1[2]), a);
	float s = glm::mix(s0, s1, a);
	quat q0, q1;
	q0.x = r0_xyzw[0], q0.y = r0_xyzw[1], q0.z = r0_xyzw[2], q0.w = r0_xyzw[3];
	q1.x = r1_xyzw[0], q1.y = r1_xyzw[1], q1.z = r1_xyzw[2], q1.w = r1_xyzw[3];
	quat q = glm::slerp(q0, q1, a);
	out8[0] = p.x, out8[1] = p.y, out8[2] = p.z, out8[3] = s;
	out8[4] = q.x, out8[5] = q.y, out8[6] = q.z, out8[7] = q.w;
}

// CullData for one camera through glm (host_golden.h fillCullData), for property tests of nvc_host_cull_data
extern "C" void cull_data_ref(const float* pos3, const float* quat_xyzw, float fovY, float znear, uint32_t width, uint32_t height, uint32_t drawCount, uint32_t lodStep, void* out144)
{
	CameraCase c = { { pos3[0], pos3[1], pos3[2] }, { quat_xyzw[0], quat_xyzw[1], quat_xyzw[2], quat_xyzw[3] }, fovY, znear, width, height, drawCount, lodStep };
	CullData cd = fillCullData(c);
	memcpy(out144, &cd, sizeof(cd));
}
