// TEST INFRASTRUCTURE: VX3_MemoryCleaner.h guards a vector with a boost::mutex.
#pragma once
#include <mutex>
namespace boost {
using mutex = std::mutex;
}
