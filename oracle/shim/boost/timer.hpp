// Test-infrastructure stand-in for "boost/timer.hpp" (CPU-clock timer, as boost::timer v1).
#pragma once
#include <ctime>
namespace boost {
class timer
{
public:
    timer() : _start(std::clock()) {}
    void restart() { _start = std::clock(); }
    double elapsed() const { return double(std::clock() - _start) / CLOCKS_PER_SEC; }

private:
    std::clock_t _start;
};
} // namespace boost
