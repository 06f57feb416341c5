// Stand-in for the FleCSI runtime entry points the reference drivers use
// (app/drivers/hydro/main.cc:47, main_driver.cc:282,309).
#pragma once
namespace flecsi { namespace execution {
void specialization_tlt_init(int argc, char ** argv);
void driver(int argc, char ** argv);
struct context_t {
  static context_t & instance() { static context_t c; return c; }
  int initialize(int argc, char ** argv) {
    specialization_tlt_init(argc, argv);
    driver(argc, argv);
    return 0;
  }
};
} }
#define flecsi_register_mpi_task(task, nspace) static_assert(true, "")
#define flecsi_execute_mpi_task(task, nspace, ...) nspace::task(__VA_ARGS__)
