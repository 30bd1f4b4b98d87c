// Empty stand-in so the reference's compressor headers (which include
// <tbb/parallel_for.h> but use nothing from it) compile in an image without TBB.
#pragma once
#include <algorithm>
#include <cstdint>
