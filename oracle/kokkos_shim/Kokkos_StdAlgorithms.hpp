// TEST INFRASTRUCTURE ONLY -- stand-in header; the reference only needs
// Kokkos::Experimental::swap from here, which Kokkos_Core.hpp of this shim provides.
#pragma once
#include "Kokkos_Core.hpp"
