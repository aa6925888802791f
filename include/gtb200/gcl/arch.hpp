/*
 * gtb200/gcl/arch.hpp -- the architecture tag of the B200 halo exchange and boundary conditions, next to
 * gridtools::gcl::cpu and gridtools::gcl::gpu (gcl/low_level/arch.hpp:28,32).
 */
#pragma once

namespace gridtools {
    namespace gcl {
        /// Indicates that the data lives on a B200 and is handled by libgtb200 / the b200 headers.
        struct b200 {};
    } // namespace gcl
} // namespace gridtools
