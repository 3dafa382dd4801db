// Thread-local error message behind pdx_last_error() (defined in pdx_abi.cu), shared by every
// translation unit that implements extern "C" entry points.
#pragma once
namespace pdx {
// stores `msg` for pdx_last_error() and returns `code`
int set_error(int code, const char* msg);
}
