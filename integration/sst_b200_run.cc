// Run-mode hand-off from stock SST-core to the device engine:
//
//     sst -n 2 --partitioner=b200.run [--stop-at T] [--model-options "..."] model.py   (library libb200.so on --lib-path)
//
// (-n 2: a serial stock-SST job never consults its partitioner, main.cc:1063-1065; the thread count plays no other role,
// the hand-off happens before any thread is started.)
//
// Stock SST parses the model script with its own embedded Python, builds and checks its ConfigGraph, and then asks the
// partitioner plugin for an assignment (main.cc:388-438).  A partitioner that requires the ConfigGraph receives the whole
// graph (sstpart.h:60-62) -- every component with its type and Params, every link with its ports and latencies, the
// statistics configuration -- which is everything sstb200_engine_create's driver needs.  This plugin therefore
//   1. writes the graph with the core's OWN interchange writer (model/cfgoutput/jsonConfigOutput.cc, the code behind
//      --output-json), together with the run's Config (stop-at, time base: Simulation::config, simulation.h:713);
//   2. hands that file to the device engine's host driver inside the interpreter stock SST already embeds
//      (sst_core_b200.run: ConfigGraph -> wired arrays -> sstb200_engine_create / setup / run / results, include/sstb200.h);
//      the driver prints what Simulation::run + finish() would have printed;
//   3. ends the process with the driver's exit code: Simulation::run never starts on the host.
// Models with an element the device engine does not have are refused with the driver's message and exit code 1 (the
// reference's "can't find requested component" fatal).  tests/test_integration_shim.py runs the unmodified reference
// binary with this plugin.
#include "sst_config.h"

#include "sst/core/sstpart.h"

#include "sst/core/eli/elementinfo.h"
#include "sst/core/model/cfgoutput/jsonConfigOutput.h"
#include "sst/core/model/configGraph.h"
#include "sst/core/output.h"
#include "sst/core/rankInfo.h"
#include "sst/core/simulation.h"

#include <Python.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

namespace B200 {

class RunOnDevice : public SST::Partition::SSTPartitioner
{
public:
    SST_ELI_REGISTER_PARTITIONER(
        RunOnDevice,
        "b200",
        "run",
        SST_ELI_ELEMENT_VERSION(1,0,0),
        "hands the ConfigGraph to the B200 device engine (libsstb200) and ends the run with its results")

    RunOnDevice(SST::RankInfo total, SST::RankInfo /*my_rank*/, int /*verbosity*/) : total_(total) {}

    bool requiresConfigGraph() override { return true; }
    bool spawnOnAllRanks() override { return false; }

    void performPartition(SST::ConfigGraph* graph) override
    {
        SST::Output& out = SST::Output::getDefaultObject();
        if ( total_.rank != 1 )
            out.fatal(CALL_INFO, 1, "b200.run: run stock SST as one MPI rank; the device engine partitions the model over "
                                    "the GPUs itself\n");
        // 1. the graph + the run's Config, in the core's interchange format
        char path[] = "/tmp/sstb200_graph_XXXXXX.json";
        int  fd     = mkstemps(path, 5);
        if ( fd < 0 ) out.fatal(CALL_INFO, 1, "b200.run: cannot create a temporary file\n");
        close(fd);
        {
            SST::Core::JSONConfigGraphOutput writer(path);
            writer.generate(&SST::Simulation::config, graph);
        }
        // 2. the device engine's host driver, inside the interpreter that read the model script
        const char* root = getenv("SSTB200_ROOT"); // where the sst_core_b200 package lives
        std::string code = "import sys\n";
        if ( root ) code += std::string("sys.path.insert(0, r'") + root + "')\n";
        code += "from sst_core_b200.run import main as _sstb200_main\n";
        code += std::string("_sstb200_rc = _sstb200_main([r'") + path + "'])\n";
        fflush(stdout);
        PyGILState_STATE gil = PyGILState_Ensure();
        int              rc  = PyRun_SimpleString(code.c_str());
        long             ret = 1;
        if ( rc == 0 ) {
            PyObject* mainmod = PyImport_AddModule("__main__");
            PyObject* v       = mainmod ? PyObject_GetAttrString(mainmod, "_sstb200_rc") : nullptr;
            if ( v ) {
                ret = PyLong_AsLong(v);
                Py_DECREF(v);
            }
            PyRun_SimpleString("import sys; sys.stdout.flush(); sys.stderr.flush()");
        }
        PyGILState_Release(gil);
        unlink(path);
        // 3. the run is over: Simulation::run never starts on the host
        fflush(stdout);
        _exit((int)ret);
    }

private:
    SST::RankInfo total_;
};

} // namespace B200
