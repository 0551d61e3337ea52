// TEST INFRASTRUCTURE ONLY.  Stand-in for the reference's src/io/curl.cxx (libcurl is not in this image).
// Only the `download` subcommand calls these (reference src/guidescan.cxx:260-314); `index` and
// `enumerate`, the paths the oracle exists for, never do.  Signatures: reference include/io/curl.hpp:6-7.
#include <string>
#include "io/curl.hpp"
namespace io {
int download_file(std::string, std::string) { return 1; }
int download_json(std::string, json&) { return 1; }
}
