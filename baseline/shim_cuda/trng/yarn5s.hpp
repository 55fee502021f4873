#pragma once
#include "philox_engine.hpp"
