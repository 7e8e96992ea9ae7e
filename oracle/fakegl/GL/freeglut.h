/* GL/freeglut.h -- TEST INFRASTRUCTURE (oracle).  Stand-in for freeglut so that the
 * reference's horizonator-lib.c compiles unmodified; the functions are no-ops in fakegl.c
 * (there is no window system). */
#pragma once
#define GLUT_RGB 0x0000
#define GLUT_DOUBLE 0x0002
#define GLUT_DEPTH 0x0010
#define GLUT_CORE_PROFILE 0x0001
#define GLUT_FORWARD_COMPATIBLE 0x0002

void glutInitContextFlags(int flags);
void glutInitContextVersion(int major, int minor);
void glutInitContextProfile(int profile);
void glutInit(int* argc, char** argv);
void glutInitDisplayMode(unsigned int mode);
void glutInitWindowSize(int w, int h);
int  glutCreateWindow(const char* title);
void glutHideWindow(void);
int  glutExtensionSupported(const char* ext);
void glutSetWindow(int id);
void glutDestroyWindow(int id);
void glutExit(void);
