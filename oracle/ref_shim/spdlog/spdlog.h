// Minimal spdlog stand-in — TEST INFRASTRUCTURE (oracle/_ref build only).
#pragma once
#include <fmt/format.h>
#include <cstdlib>
#define SPDLOG_INFO(...) do { if (std::getenv("REF_VERBOSE")) { fmt::print(stderr, "[info] "); fmt::print(stderr, __VA_ARGS__); fmt::print(stderr, "\n"); } } while (0)
#define SPDLOG_CRITICAL(...) do { fmt::print(stderr, "[critical] "); fmt::print(stderr, __VA_ARGS__); fmt::print(stderr, "\n"); } while (0)
