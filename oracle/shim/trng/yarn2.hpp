#pragma once
#include "feed.hpp"
namespace trng { struct yarn2 : public feed_engine {}; }
