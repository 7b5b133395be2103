// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- stand-in for meta/include/gnuradio-4.0/meta/formatter.hpp
// (needs GCC >= 14); the formatters the compiled reference sources need live in the Message.hpp stand-in.
#ifndef GR4B200_ORACLE_SHIM_FORMATTER_HPP
#define GR4B200_ORACLE_SHIM_FORMATTER_HPP
#include <gnuradio-4.0/Message.hpp>
#endif
