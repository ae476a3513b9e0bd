// cuam_pybind.cpp -- pybind11 module `cuam`: the reference's extension module name, its five
// functions and their keyword arguments (reference backend/src/cuam.cpp:186-217), implemented on
// top of the C ABI in include/am_b200.h.  No libtorch: tensors are duck-typed through
// data_ptr()/shape/dtype/is_cuda/is_contiguous() (torch.Tensor) or __array_interface__ (numpy),
// so the module builds in seconds and works for host and device tensors alike.
//
// Argument checks mirror the reference's TORCH_CHECKs (cuam.cpp:58-184) and raise RuntimeError;
// call-order violations print the reference's messages and return (cuam_kernel.cu:172-176,
// 193-197, 241-245); CUDA failures raise instead of exit()-ing.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "am_b200.h"

namespace py = pybind11;

namespace {

am_handle *g_handle = nullptr;          // process-global environment, like the reference's var_ptr<T>
std::string g_float_type;
std::vector<int> g_nodes;
std::vector<int> g_arc;                 // flattened arc_table
int g_arc_rows = 0, g_arc_cols = 0;

void check(bool ok, const std::string &what)
{
    if (!ok) throw std::runtime_error("Expected " + what + " to be true, but got false.");
}

struct View {
    void *ptr = nullptr;
    std::vector<int64_t> shape;
    std::string dtype;
    bool cuda = false;
};

View view_of(const py::handle &t, const char *name)
{
    View v;
    if (py::hasattr(t, "data_ptr")) {   // torch.Tensor
        check(t.attr("is_contiguous")().cast<bool>(), std::string(name) + " must be contiguous");
        v.cuda = t.attr("is_cuda").cast<bool>();
        for (auto s : t.attr("shape")) v.shape.push_back(s.cast<int64_t>());
        v.dtype = py::str(t.attr("dtype")).cast<std::string>();
        if (v.dtype.rfind("torch.", 0) == 0) v.dtype = v.dtype.substr(6);
        v.ptr = t.attr("numel")().cast<int64_t>() ? reinterpret_cast<void *>(t.attr("data_ptr")().cast<uintptr_t>())
                                                  : nullptr;
    } else {                            // numpy.ndarray
        check(t.attr("flags").attr("c_contiguous").cast<bool>(), std::string(name) + " must be contiguous");
        for (auto s : t.attr("shape")) v.shape.push_back(s.cast<int64_t>());
        v.dtype = py::str(t.attr("dtype")).cast<std::string>();
        v.ptr = t.attr("size").cast<int64_t>() ? reinterpret_cast<void *>(t.attr("ctypes").attr("data").cast<uintptr_t>())
                                               : nullptr;
    }
    return v;
}

void fail_with(const char *what, int rc)
{
    const char *m = am_last_error(g_handle);
    throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + (m ? m : ""));
}

void Destroy()
{
    if (!g_handle) {
        std::cout << "Environment must be initialized first!" << std::endl;
        return;
    }
    am_destroy(g_handle);
    g_handle = nullptr;
    g_float_type.clear();
}

void Init(const std::string &float_type, const std::vector<int> &nodesnum, const py::object &arc_table,
          int num_extra_constraints)
{
    if (float_type != "float32" && float_type != "float64") {
        std::cout << "Error: `float_type` is either `float32` or `float64`!";
        return;
    }
    check(!nodesnum.empty() && nodesnum.front() == 3, "nodesnum.front() == 3");
    check(nodesnum.back() == 1, "nodesnum.back() == 1");
    check(nodesnum.size() >= 3, "nodesnum.size() >= 3");
    View a = view_of(arc_table, "arc_table");
    check(!a.cuda, "arc_table must be a CPU tensor");
    check(a.shape.size() == 2, "arc_table.dim() == 2");
    check(a.shape[0] == (int64_t)nodesnum.size() - 2, "arc_table.size(0) == nodesnum.size() - 2");
    check(a.shape[1] >= 1 && (a.shape[1] + 1) % 2 == 0, "arc_table.size(1) >= 1 && (arc_table.size(1) + 1) % 2 == 0");
    check(a.dtype == "int32", "arc_table.dtype() == torch::kInt32");
    const int *at = static_cast<const int *>(a.ptr);
    for (int64_t i = 0; i < a.shape[0]; ++i) check(1 + 2 * at[i * a.shape[1]] <= a.shape[1], "1 + 2 * connection_num <= arc_table.size(1)");
    check(num_extra_constraints >= 0, "num_extra_constraints >= 0");
    if (g_handle) Destroy();
    am_handle *h = nullptr;
    const int rc = am_create(&h, float_type == "float64", nodesnum.data(), (int)nodesnum.size(), at, (int)a.shape[0],
                             (int)a.shape[1], num_extra_constraints);
    if (rc != AM_OK) throw std::runtime_error(std::string("Init failed: ") + am_last_error(nullptr));
    g_handle = h;
    g_float_type = float_type;
    g_nodes = nodesnum;
    g_arc.assign(at, at + a.shape[0] * a.shape[1]);
    g_arc_rows = (int)a.shape[0];
    g_arc_cols = (int)a.shape[1];
}

void AnalyticMarching(const py::list &weights, const py::list &biases, const py::object &states, const py::object &points,
                      const py::list &arc_tm, const py::object &w_extra_constraints, const py::object &b_extra_constraints,
                      double iso, bool flip_insideout)
{
    if (g_float_type.empty()) {
        std::cout << "Environment must be initialized first!" << std::endl;
        return;
    }
    const std::string &want = g_float_type;
    check(weights.size() == biases.size(), "weights.size() == biases.size()");
    const int n_fc = (int)weights.size();
    check(n_fc == (int)g_nodes.size() - 1, "fc_layers_num == nodesnum_.size() - 1");
    std::vector<View> W, B, TM;
    int64_t L = 0;
    for (int i = 0; i < n_fc; ++i) {
        W.push_back(view_of(weights[i], "weights[i]"));
        B.push_back(view_of(biases[i], "biases[i]"));
        check(W[i].shape.size() == 2, "weights[i].dim() == 2");
        check(B[i].shape.size() == 1, "biases[i].dim() == 1");
        check(W[i].shape[0] == B[i].shape[0], "weights[i].size(0) == biases[i].size(0)");
        if (i) check(W[i].shape[1] == W[i - 1].shape[0], "weights[i].size(1) == weights[i - 1].size(0)");
        L += W[i].shape[0];
        check(W[i].dtype == want && B[i].dtype == want, "weights[i].dtype() == biases[i].dtype() == " + want);
        check(g_nodes[i + 1] == W[i].shape[0], "nodesnum_[i + 1] == weights[i].size(0)");
    }
    check(W[0].shape[1] == 3, "weights[0].size(1) == 3");
    check(W.back().shape[0] == 1, "weights.back().size(0) == 1");
    L -= 1;
    View S = view_of(states, "states"), P = view_of(points, "points");
    check(S.shape.size() == 2 && S.shape[0] >= 1 && S.shape[1] == L, "states.size() == (N >= 1, hidden_states_vector_len)");
    check(S.dtype == "bool", "states.dtype() == torch::kBool");
    check(P.shape.size() == 2 && P.shape[0] == S.shape[0] && P.shape[1] == 3, "points.size() == (N, 3)");
    check(P.dtype == want, "points.dtype() == " + want);
    std::vector<int> tm_shapes;
    for (auto t : arc_tm) {
        TM.push_back(view_of(t, "arc_tm[i]"));
        check(TM.back().shape.size() == 2, "tm.dim() == 2");
        check(TM.back().dtype == want, "tm.dtype() == " + want);
        tm_shapes.push_back((int)TM.back().shape[0]);
        tm_shapes.push_back((int)TM.back().shape[1]);
    }
    for (int i = 0; i < g_arc_rows; ++i)
        for (int j = 0; j < g_arc[i * g_arc_cols]; ++j) {
            const int from = g_arc[i * g_arc_cols + 2 * j + 1], idx = g_arc[i * g_arc_cols + 2 * j + 2];
            check(idx < (int)TM.size(), "arc_table references a transform that was not given");
            if (TM[idx].shape[0] || TM[idx].shape[1]) {
                check(TM[idx].shape[0] == W[i + 1].shape[0], "arc_tm[idx].size(0) == weights[i + 1].size(0)");
                check(TM[idx].shape[1] == W[from].shape[1], "arc_tm[idx].size(1) == weights[from].size(1)");
            }
        }
    View WE = view_of(w_extra_constraints, "w_extra_constraints"), BE = view_of(b_extra_constraints, "b_extra_constraints");
    check(WE.shape.size() == 2 && WE.shape[1] == 3, "w_extra_constraints.size(1) == 3");
    check(BE.shape.size() == 1 && BE.shape[0] == WE.shape[0], "w_extra_constraints.size(0) == b_extra_constraints.size(0)");
    check(WE.dtype == want && BE.dtype == want, "extra constraints dtype == " + want);
    std::vector<const void *> wp, bp, tp;
    for (auto &v : W) wp.push_back(v.ptr);
    for (auto &v : B) bp.push_back(v.ptr);
    for (auto &v : TM) tp.push_back(v.ptr);
    if (tp.empty()) tp.push_back(nullptr);
    if (tm_shapes.empty()) tm_shapes.push_back(0);
    int rc;
    {
        py::gil_scoped_release nogil;   // the reference holds the GIL for the whole march
        rc = am_march(g_handle, wp.data(), bp.data(), tp.data(), tm_shapes.data(), (int)TM.size(),
                      static_cast<const uint8_t *>(S.ptr), P.ptr, S.shape[0], WE.ptr, BE.ptr, (int)WE.shape[0], iso,
                      flip_insideout ? 1 : 0, nullptr);
    }
    if (rc != AM_OK) fail_with("AnalyticMarching", rc);
}

void CombineMesh(double scale, const std::vector<double> &center)
{
    check(center.size() == 3, "center.size() == 3");
    const int rc = g_handle ? am_combine(g_handle, scale, center.data()) : AM_ERR_STATE;
    if (rc == AM_ERR_STATE) {
        std::cout << "AnalyticMarching must be done first!" << std::endl;
        return;
    }
    if (rc != AM_OK) fail_with("CombineMesh", rc);
}

void ExportMesh(const std::string &file_path, bool is_polymesh, bool is_float32)
{
    const int rc = g_handle ? am_export(g_handle, file_path.c_str(), is_polymesh, is_float32) : AM_ERR_STATE;
    if (rc == AM_ERR_STATE) {
        std::cout << "CombineMesh must be done first!" << std::endl;
        return;
    }
    if (rc != AM_OK) fail_with("ExportMesh", rc);
}

py::dict Stats()
{
    am_stats s{};
    if (!g_handle || am_get_stats(g_handle, &s) != AM_OK) throw std::runtime_error("no environment");
    py::dict d;
    d["n_seeds"] = s.n_seeds; d["n_unique_seeds"] = s.n_unique_seeds; d["n_states"] = s.n_states;
    d["n_faces"] = s.n_faces; d["n_corners"] = s.n_corners; d["n_levels"] = s.n_levels;
    d["n_candidates"] = s.n_candidates; d["n_unbounded"] = s.n_unbounded; d["n_overflow"] = s.n_overflow;
    d["n_over_vertmax"] = s.n_over_vertmax; d["n_inconsistent"] = s.n_inconsistent; d["n_vertices"] = s.n_vertices;
    d["n_stitch_miss"] = s.n_stitch_miss; d["max_level_states"] = s.max_level_states; d["n_launches"] = s.n_launches;
    d["seconds_march"] = s.seconds_march; d["seconds_compose"] = s.seconds_compose; d["seconds_clip"] = s.seconds_clip;
    d["seconds_frontier"] = s.seconds_frontier; d["compose_flops"] = s.compose_flops;
    return d;
}

}  // namespace

PYBIND11_MODULE(cuam, m)
{
    m.doc() = "B200-native implementation of the Analytic Marching algorithm (drop-in for AnalyticMesh's cuam)";
    m.def("Init", &Init, "Initialize environment (CUDA)", py::arg("float_type"), py::arg("nodesnum"), py::arg("arc_table"),
          py::arg("num_extra_constraints"));
    m.def("AnalyticMarching", &AnalyticMarching, "AnalyticMarching (CUDA)", py::arg("weights"), py::arg("biases"),
          py::arg("states"), py::arg("points"), py::arg("arc_tm"), py::arg("w_extra_constraints"),
          py::arg("b_extra_constraints"), py::arg("iso"), py::arg("flip_insideout"));
    m.def("CombineMesh", &CombineMesh, "Combine to a mesh (CUDA)", py::arg("scale"), py::arg("center"));
    m.def("ExportMesh", &ExportMesh, "Export mesh file (CUDA)", py::arg("file_path"), py::arg("is_polymesh"),
          py::arg("is_float32"));
    m.def("Destroy", &Destroy, "Destroy environment (CUDA)");
    m.def("Stats", &Stats, "Counters of the last march (not in the reference)");
}
