#pragma once
#include <filesystem>
namespace boost { namespace filesystem {
using std::filesystem::path; using std::filesystem::exists; using std::filesystem::is_directory;
using std::filesystem::rename; using std::filesystem::create_directories;
using std::filesystem::directory_iterator; using std::filesystem::remove;
} }
