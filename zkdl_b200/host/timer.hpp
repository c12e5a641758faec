// timer.hpp — drop-in forwarding header (Timer lives in zkdl.hpp).
#pragma once
#include "zkdl.hpp"
