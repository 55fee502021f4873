#pragma once
#include "yarn2.hpp"
namespace trng { typedef yarn2 yarn5s; }
