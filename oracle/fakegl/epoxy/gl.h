/* epoxy/gl.h -- TEST INFRASTRUCTURE (oracle).  Stand-in for libepoxy's header so that the
 * reference's horizonator-lib.c compiles UNMODIFIED from /root/reference without any GL
 * installed.  Declares exactly the types, enums and entry points that file uses; they are
 * implemented by fakegl.c on top of the GL-pipeline restatement (oracle/gl_pipeline.c). */
#pragma once
#include <stddef.h>
#include <stdint.h>

typedef int            GLint;
typedef unsigned int   GLuint;
typedef short          GLshort;
typedef char           GLchar;
typedef void           GLvoid;
typedef unsigned int   GLenum;
typedef int            GLsizei;
typedef float          GLfloat;
typedef unsigned char  GLboolean;
typedef unsigned int   GLbitfield;
typedef ptrdiff_t      GLsizeiptr;

/* enum values are those of the Khronos registry (gl.xml) */
#define GL_FALSE 0
#define GL_TRUE  1
#define GL_NO_ERROR 0
#define GL_TRIANGLES 0x0004
#define GL_DEPTH_BUFFER_BIT 0x00000100
#define GL_COLOR_BUFFER_BIT 0x00004000
#define GL_CULL_FACE  0x0B44
#define GL_DEPTH_TEST 0x0B71
#define GL_VIEWPORT   0x0BA2
#define GL_PACK_ALIGNMENT 0x0D05
#define GL_TEXTURE_2D 0x0DE1
#define GL_SHORT 0x1402
#define GL_UNSIGNED_BYTE 0x1401
#define GL_UNSIGNED_INT 0x1405
#define GL_FLOAT 0x1406
#define GL_DEPTH_COMPONENT 0x1902
#define GL_RGB 0x1907
#define GL_BGR 0x80E0
#define GL_VERSION 0x1F02
#define GL_LINEAR 0x2601
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_WRAP_S 0x2802
#define GL_TEXTURE_WRAP_T 0x2803
#define GL_REPEAT 0x2901
#define GL_TEXTURE0_ARB 0x84C0
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_WRITE_ONLY 0x88B9
#define GL_STATIC_DRAW 0x88E4
#define GL_FRAGMENT_SHADER 0x8B30
#define GL_VERTEX_SHADER 0x8B31
#define GL_GEOMETRY_SHADER 0x8DD9
#define GL_FRAMEBUFFER_COMPLETE 0x8CD5
#define GL_COLOR_ATTACHMENT0 0x8CE0
#define GL_DEPTH_ATTACHMENT 0x8D00
#define GL_FRAMEBUFFER 0x8D40
#define GL_RENDERBUFFER 0x8D41

GLenum glGetError(void);
const unsigned char* glGetString(GLenum name);
void glEnable(GLenum cap);
void glClearColor(GLfloat r, GLfloat g, GLfloat b, GLfloat a);
void glClear(GLbitfield mask);
void glViewport(GLint x, GLint y, GLsizei w, GLsizei h);
void glGetIntegerv(GLenum pname, GLint* data);
void glPixelStorei(GLenum pname, GLint param);
void glDrawBuffer(GLenum buf);
void glReadPixels(GLint x, GLint y, GLsizei w, GLsizei h, GLenum format, GLenum type, void* pixels);
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const void* indices);

void glGenVertexArrays(GLsizei n, GLuint* ids);
void glBindVertexArray(GLuint id);
void glGenBuffers(GLsizei n, GLuint* ids);
void glBindBuffer(GLenum target, GLuint id);
void glBufferData(GLenum target, GLsizeiptr size, const void* data, GLenum usage);
void* glMapBuffer(GLenum target, GLenum access);
GLboolean glUnmapBuffer(GLenum target);
void glEnableVertexAttribArray(GLuint index);
void glVertexAttribPointer(GLuint index, GLint size, GLenum type, GLboolean normalized,
                           GLsizei stride, const void* pointer);

GLuint glCreateProgram(void);
GLuint glCreateShader(GLenum type);
void glShaderSource(GLuint shader, GLsizei count, const GLchar** string, const GLint* length);
void glCompileShader(GLuint shader);
void glGetShaderInfoLog(GLuint shader, GLsizei bufsize, GLsizei* length, GLchar* log);
void glAttachShader(GLuint program, GLuint shader);
void glLinkProgram(GLuint program);
void glUseProgram(GLuint program);
void glGetProgramInfoLog(GLuint program, GLsizei bufsize, GLsizei* length, GLchar* log);
GLint glGetUniformLocation(GLuint program, const GLchar* name);
void glUniform1f(GLint location, GLfloat v);
void glUniform1i(GLint location, GLint v);
void glGetUniformfv(GLuint program, GLint location, GLfloat* params);

void glGenFramebuffers(GLsizei n, GLuint* ids);
void glBindFramebuffer(GLenum target, GLuint id);
void glGenRenderbuffers(GLsizei n, GLuint* ids);
void glBindRenderbuffer(GLenum target, GLuint id);
void glRenderbufferStorage(GLenum target, GLenum internalformat, GLsizei w, GLsizei h);
void glFramebufferRenderbuffer(GLenum target, GLenum attachment, GLenum rbtarget, GLuint rb);
GLenum glCheckFramebufferStatus(GLenum target);

/* texture path: declared so the file compiles; never reached on the untextured path */
void glGenTextures(GLsizei n, GLuint* ids);
void glActiveTextureARB(GLenum texture);
void glBindTexture(GLenum target, GLuint id);
void glTexParameteri(GLenum target, GLenum pname, GLint param);
void glTexImage2D(GLenum target, GLint level, GLint internalformat, GLsizei w, GLsizei h,
                  GLint border, GLenum format, GLenum type, const void* pixels);
void glTexSubImage2D(GLenum target, GLint level, GLint xo, GLint yo, GLsizei w, GLsizei h,
                     GLenum format, GLenum type, const void* pixels);

/* not GL: knobs of the fake driver */
void fakegl_set_threads(int nthreads);
