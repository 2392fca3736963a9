// Error taxonomy of the facade — same names and hierarchy as the reference (src/CadR/Exceptions.h:13-40) so that
// `catch(CadR::Error&)` in applications (examples/RenderingPerformance/main.cpp:1803) keeps working.  The C ABI
// never throws; cadr::check() turns its error codes into these exceptions.
#pragma once
#include <stdexcept>
#include <string>

namespace CadR {

class Error {
public:
	virtual ~Error() = default;
	virtual const char* what() const noexcept = 0;
};

class LogicError : public Error, public std::logic_error {
public:
	explicit LogicError(const std::string& w) : std::logic_error(w) {}
	const char* what() const noexcept override { return std::logic_error::what(); }
};

class OutOfResources : public Error, public std::runtime_error {
public:
	explicit OutOfResources(const std::string& w) : std::runtime_error(w) {}
	const char* what() const noexcept override { return std::runtime_error::what(); }
};

class Timeout : public Error, public std::runtime_error {
public:
	explicit Timeout(const std::string& w) : std::runtime_error(w) {}
	const char* what() const noexcept override { return std::runtime_error::what(); }
};

// CUDA runtime/driver failure reported by the backend (the reference throws vk::Error here)
class DeviceError : public Error, public std::runtime_error {
public:
	explicit DeviceError(const std::string& w) : std::runtime_error(w) {}
	const char* what() const noexcept override { return std::runtime_error::what(); }
};

void check(int code);  // throws the exception matching a cadr_b200 error code

}
