// rs_shader.h — descriptor of one generated shader TU (see gen.py).  TEST INFRASTRUCTURE ONLY (oracle/).
#pragma once
#include <stddef.h>

struct RsShader
{
	const char* source;                               // file name under src/shaders
	const unsigned int* local_size;                   // layout(local_size_x/y/z)
	void (*main)();                                   // the shader's main(), one invocation
	void (*bind)(int binding, void* p, size_t bytes); // descriptor write
	void (*push)(const void* p, size_t bytes);        // push constants
	void (*spec)(int constant_id, int value);         // specialisation constants
	void* (*payload)();                               // taskPayloadSharedEXT object of the calling host thread, or NULL
	void* (*output)(int location);                    // layout(location = N) out array of the calling host thread, or NULL
	int uses_barrier;
};
