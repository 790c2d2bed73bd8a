/* Minimal <GL/gl.h> stand-in: the three names cuda_gl_interop.h and hg_context.cu need.
 *
 * The build image has no OpenGL headers or libraries, so the CUDA-GL interop path of
 * libhydrogen_b200.so (-DHG_WITH_GL, hg_register_gl / hg_publish_gl) cannot be linked or run here.  This file lets
 * that path be COMPILED and type-checked (`make -C hydro_gen_b200/csrc gl-syntax`): a real build puts the system's
 * GL include directory first and never sees it.  Values are those of the OpenGL registry (gl.xml). */
#ifndef HG_GL_STUB_H
#define HG_GL_STUB_H
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
#define GL_TEXTURE_2D 0x0DE1
#define GL_RGBA32F 0x8814
#define GL_VERSION 0x1F02
#endif
