#pragma once
#include "feed.hpp"
namespace trng { struct yarn5s : public feed_engine {}; }
