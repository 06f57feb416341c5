// Empty stand-in for FleCSI's data client base (reference: tree_topology.h:46,69).
#pragma once
namespace flecsi { namespace data { struct data_client_t {}; } }
