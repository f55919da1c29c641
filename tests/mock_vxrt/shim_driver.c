/*
 * shim_driver.c -- TEST INFRASTRUCTURE (CPU).  Calls the GL entry points of libvxrt_glshim.so the way a GL program would
 * (through GLEW's function-pointer variables) on a tiny 8x4x8 grid, to check that batching glBufferSubData calls keeps
 * GL's ordering semantics: later writes to the same voxels win, data is captured at call time, a whole-buffer
 * glBufferData supersedes what was pending.  tests/test_glshim.py compares the mock's final grid with its own replay.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned int GLenum, GLuint;
typedef int GLint, GLsizei;
typedef char GLchar;
extern GLuint (*__glewCreateShader)(GLenum);
extern void (*__glewShaderSource)(GLuint, GLsizei, const GLchar* const*, const GLint*);
extern void (*__glewBufferData)(GLenum, ptrdiff_t, const void*, GLenum);
extern void (*__glewBufferSubData)(GLenum, ptrdiff_t, ptrdiff_t, const void*);
extern GLint (*__glewGetUniformLocation)(GLuint, const GLchar*);
extern void (*__glewUniform3f)(GLint, float, float, float);
extern void glDrawArrays(GLenum, GLint, GLsizei);
extern void glViewport(GLint, GLint, GLsizei, GLsizei);
extern void glfwSwapBuffers(void*);
extern int glfwWindowShouldClose(void*);

#define SSBO 0x90D2
#define N (8 * 4 * 8)

int main(int argc, char** argv) {
    const char* src = argc > 1 ? argv[1] : "const int VOXELS_WIDTH=8;\nconst int VOXELS_HEIGHT=4;\nconst int RENDER_DIST=384;\n";
    GLuint fs = __glewCreateShader(0x8B30);
    __glewShaderSource(fs, 1, &src, NULL);
    int v[N], w[N];
    for (int i = 0; i < N; i++) { v[i] = i; w[i] = 1000 + i; }
    glViewport(0, 0, 64, 32);
    __glewBufferData(SSBO, sizeof v, v, 0x88EA);
    __glewUniform3f(__glewGetUniformLocation(1, "camPos"), 1.f, 2.f, 3.f);
    /* batch 1: three rows of 4 */
    __glewBufferSubData(SSBO, 4 * 8, 16, &w[8]);
    __glewBufferSubData(SSBO, 4 * 40, 16, &w[40]);
    __glewBufferSubData(SSBO, 4 * 100, 16, &w[100]);
    w[8] = -5;                                          /* after the call: must NOT reach the buffer */
    /* overlaps the first row with other data: a later write wins */
    int late[4] = {7001, 7002, 7003, 7004};
    __glewBufferSubData(SSBO, 4 * 10, 16, late);
    /* another length */
    __glewBufferSubData(SSBO, 4 * 200, 8, &w[200]);
    glDrawArrays(4, 0, 6);
    glfwSwapBuffers(NULL);
    if (glfwWindowShouldClose(NULL)) return 3;          /* VXRT_GLSHIM_FRAMES=2: not yet */
    /* pending rows are dropped by a whole-buffer respecification */
    __glewBufferSubData(SSBO, 4 * 50, 16, &w[60]);
    for (int i = 0; i < N; i++) v[i] = 2 * i;
    v[8] = 16; v[9] = 18;
    __glewBufferData(SSBO, sizeof v, v, 0x88EA);
    __glewBufferSubData(SSBO, 4 * 252, 16, &w[252]);     /* the last row of the buffer */
    glDrawArrays(4, 0, 6);
    glfwSwapBuffers(NULL);
    return glfwWindowShouldClose(NULL) ? 0 : 4;         /* two frames presented: the window closes, the context goes */
}
