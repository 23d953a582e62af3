// yk/config.hpp -- the reference's global verbosity switch (yk/config.hpp:28-30).
#pragma once
#include <cstdint>
namespace yk {
inline std::uint32_t verbose = 0;
inline bool stats = false;          // print the kernel timing line(s) without the reference's per-pixel progress lines (raytrace --stats)
}
