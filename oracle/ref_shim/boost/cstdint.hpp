#pragma once
#include <stdint.h>
namespace boost { using ::uint8_t; }
