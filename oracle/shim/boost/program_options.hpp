// Test-infrastructure stand-in for <boost/program_options.hpp>.
//
// Boost is not installed in this image. The reference only touches Boost in its CLI
// option-binding code (src/configurations/*.cpp), which the oracle harness never calls:
// config structs are filled programmatically, the way test/test.cpp:170-179 does.
// This header supplies just enough surface for those translation units to compile.
// It carries no arithmetic and is NOT part of the product.
#pragma once
#include <algorithm>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

struct error : public std::logic_error
{
    explicit error(std::string const& what) : std::logic_error(what) {}
};

struct value_semantic
{
};

template<typename T>
struct typed_value : public value_semantic
{
    template<typename U>
    typed_value* default_value(U const&)
    {
        return this;
    }
    template<typename U>
    typed_value* default_value(U const&, std::string const&)
    {
        return this;
    }
};

template<typename T>
typed_value<T>* value(T* = nullptr)
{
    static typed_value<T> sink;
    return &sink;
}

inline typed_value<bool>* bool_switch(bool* = nullptr)
{
    static typed_value<bool> sink;
    return &sink;
}

struct options_description_easy_init
{
    options_description_easy_init& operator()(char const*, char const*) { return *this; }
    options_description_easy_init& operator()(char const*, value_semantic const*) { return *this; }
    options_description_easy_init& operator()(char const*, value_semantic const*, char const*)
    {
        return *this;
    }
};

struct options_description
{
    options_description() = default;
    explicit options_description(std::string const&) {}
    options_description_easy_init add_options() { return {}; }
};

inline std::ostream& operator<<(std::ostream& os, options_description const&)
{
    return os;
}

}} // namespace boost::program_options
