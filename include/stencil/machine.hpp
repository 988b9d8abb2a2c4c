#pragma once
// The reference's experimental UUID machine model (machine.hpp / src/machine.cpp) is not used by the
// halo-exchange path and is out of scope here; this header exists so that
// stencil/stencil.hpp's include list keeps resolving.
