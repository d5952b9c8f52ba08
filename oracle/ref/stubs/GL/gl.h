/* Empty stand-in for <GL/gl.h>, which GLFW/glfw3.h includes by default; nothing of OpenGL is used. */
#pragma once
