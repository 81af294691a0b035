// Stub for headless builds of the reference: cuda_gl_interop.h only needs these two names.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
