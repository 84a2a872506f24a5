// box2d.h — umbrella header of the B200-native drop-in for box2d-optimized's public API.
// Same role as the reference's include/box2d/box2d.h:28-61: user code includes this one file.
#ifndef BOX2D_H
#define BOX2D_H

#include "b2g_types.h"
#include "b2g_shapes.h"
#include "b2g_world.h"

#endif
