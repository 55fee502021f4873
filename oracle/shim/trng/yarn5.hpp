#pragma once
#include "feed.hpp"
namespace trng { struct yarn5 : public feed_engine {}; }
