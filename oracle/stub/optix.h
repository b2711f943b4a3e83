// Empty stand-in so that the reference's cuda/curve.h (which only #includes <optix.h> without using
// it) can be compiled on the host by oracle/ref_crosscheck.cpp.  Not OptiX; declares nothing.
#pragma once
