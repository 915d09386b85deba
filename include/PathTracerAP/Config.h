// Config.h - compile-time knobs of the drop-in facade, same names as the reference's Config.h:4-19.
// Every value can be overridden with -D (the reference hard-codes them); MAX_DEPTH names the literal 5 of Renderer.cpp:550.
#pragma once

#ifndef EPSILON
#define EPSILON 0.005f
#endif
#ifndef FLOAT_MAX
#define FLOAT_MAX 9999999.0f
#endif
#ifndef FLOAT_MIN
#define FLOAT_MIN -9999990.0f
#endif
#ifndef GRID_X
#define GRID_X 25
#define GRID_Y 25
#define GRID_Z 25
#endif
#ifndef RESOLUTION_X
#define RESOLUTION_X 1000
#endif
#ifndef RESOLUTION_Y
#define RESOLUTION_Y 800
#endif
#ifndef ITER
#define ITER 500
#endif
#ifndef MAX_DEPTH
#define MAX_DEPTH 5
#endif
#define SAMPLESX 1
#define SAMPLESY 1
#define BASE_MODEL_SCALE 1000
