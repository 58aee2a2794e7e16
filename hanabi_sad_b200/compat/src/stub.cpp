// stub.cpp -- the extension modules literally named `rela` and `hanalearn` that the reference's create.py imports and checks
// (pyhanabi/create.py:17-21 asserts `__file__.endswith(".so")`).  Compiled twice (-DHB_STUB_RELA / -DHB_STUB_HANALEARN) into
// hanabi_sad_b200/compat/{rela,hanalearn}<EXT_SUFFIX> by hanabi_sad_b200/build.py.  A stub holds no logic: it puts the repository
// root on sys.path, imports the facade module of the same name (hanabi_sad_b200.rela / .hanalearn, which drive
// libhanabi_b200.so through its C ABI) and re-exports the names the reference's pybind modules export
// (rela/pybind.cc:16-93, cpp/pybind.cc:14-56).
#include <pybind11/pybind11.h>

namespace py = pybind11;

static void export_from(py::module_& m, const char* facade, const char* const* names) {
  // <repo>/hanabi_sad_b200/compat/<this .so>  ->  <repo> on sys.path, so that `import hanabi_sad_b200` works from pyhanabi/
  py::module_ os = py::module_::import("os"), sys = py::module_::import("sys");
  py::object here = os.attr("path").attr("dirname")(os.attr("path").attr("abspath")(m.attr("__file__")));
  py::object root = os.attr("path").attr("dirname")(os.attr("path").attr("dirname")(here));
  py::list path = sys.attr("path");
  if (!path.contains(root)) path.insert(0, root);
  py::module_ impl = py::module_::import(facade);
  for (const char* const* n = names; *n; ++n) m.attr(*n) = impl.attr(*n);
  m.attr("__facade__") = impl;
}

#ifdef HB_STUB_RELA
static const char* const kNames[] = {"aggregate_priority", "BatchRunner", "Context", "FFTransition", "R2D2Actor", "RNNPrioritizedReplay",
                                     "RNNTransition", "ThreadLoop", "set_actor_duty", "set_replay_block", nullptr};
PYBIND11_MODULE(rela, m) {
  m.doc() = "drop-in for the reference's rela pybind module, served by hanabi_sad_b200.rela over libhanabi_b200.so";
  export_from(m, "hanabi_sad_b200.rela", kNames);
}
#endif

#ifdef HB_STUB_HANALEARN
static const char* const kNames[] = {"HanabiEnv", "HanabiThreadLoop", "HanabiVecEnv", nullptr};
PYBIND11_MODULE(hanalearn, m) {
  m.doc() = "drop-in for the reference's hanalearn pybind module, served by hanabi_sad_b200.hanalearn over libhanabi_b200.so";
  export_from(m, "hanabi_sad_b200.hanalearn", kNames);
}
#endif
