// TEST INFRASTRUCTURE: VX3.cuh only needs fs::path (u_with_ext); std::filesystem has the same interface.
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem = std::filesystem;
}
