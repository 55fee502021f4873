// yarn5 / yarn5s are only named by drivers outside the path; aliased so that the reference headers parse
#pragma once
#include "yarn2.hpp"
namespace trng { typedef yarn2 yarn5; }
