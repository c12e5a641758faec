// proof.cuh — drop-in forwarding header: the reference declares this part of the API in a file of this name.
#pragma once
#include "zkdl.hpp"
