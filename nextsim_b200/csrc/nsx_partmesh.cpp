// nsx_partmesh.cpp -- SURVEY.md section 8(f) row 4: the remesh-time host work in front of the path, as a fast
// host library (no GPU, no std::map / boost::bimap):
//
//   * reader of the partitioned msh-2.2 file   GmshMesh::readFromFileASCII / readFromFileBinary
//                                              (core/src/gmshmesh.cpp:133-409, 411-712)
//   * one rank's local mesh                    GmshMesh::nodalGrid (core/src/gmshmesh.cpp:856-1498),
//                                              GMSHElement::setPartition (core/include/entities.hpp:105-134)
//   * halo lists                               FiniteElement::initUpdateGhosts (model/finiteelement.cpp:14003-14088)
//   * boundary masks                           FiniteElement::bcMarkedNodes (model/finiteelement.cpp:150-271)
//   * bamg connectivity tables                 contrib/bamg/src/Mesh.cpp:526-537, 583-629, 798-865
//
// The reference derives the halo lists with MPI exchanges between the ranks; every rank reads the same file, so
// here each rank derives its neighbours' ghost sets from the file directly (O(global elements), flat arrays).
// Numbering rules (must be bit-exact): owned nodes ascending file id then ghost nodes ascending file
// id; an interface node belongs to the LOWEST rank that holds it provisionally; kept elements are the loaded ones
// with partition >= rank and three local nodes, owned elements first, file order otherwise; ghost lists per owner
// in ascending reordered id.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/nsx.h"

namespace {

thread_local std::string g_pm_err;

struct GlobalMesh {
    int nn = 0, ne = 0;
    std::vector<double> x, y;
    std::vector<int> tri;        // 3*ne, 1-based file node ids
    std::vector<int> number;     // element numbers (after the edge offset, gmshmesh.cpp:352-366)
    std::vector<int> part;       // partition % nranks
    std::vector<int> gptr, gval; // ghost partitions (CSR), % nranks
};

} // namespace

struct nsx_partmesh {
    int rank = 0, nranks = 1;
    int num_nodes = 0, local_ndof = 0, num_elements = 0, local_nelements = 0;
    int global_num_nodes = 0, global_num_elements = 0;
    std::vector<double> x, y, lat;
    std::vector<int> indices;
    std::vector<unsigned char> ghost_nodes;
    std::vector<int> node_gid, node_rid, elem_gid, elem_part, local_ghost;
    std::vector<int> send_peer, send_ptr, send_idx, recv_peer, recv_ptr, recv_idx;
    std::vector<double> nec, nc;
    int nec_w = 0, nc_w = 0;
    std::vector<unsigned char> mask_dirichlet;
    std::vector<int> neumann_flags, dirichlet_flags;
};

namespace {

// ---------------------------------------------------------------------------------------------------
// msh 2.2 reader.  Triangles only are kept; every other element type is counted as an "edge" whose count shifts
// the triangle numbers when the first triangle is not number 1 (gmshmesh.cpp:352-366, 619-634).
// ---------------------------------------------------------------------------------------------------
void swap_bytes(char* p, size_t size, size_t n)
{
    for (size_t i = 0; i < n; ++i) std::reverse(p + i * size, p + (i + 1) * size);
}

int vertices_of_type(int type)
{
    // MElement::getInfoMSH for the types Gmsh writes for a 2-D mesh
    switch (type) {
    case 1: return 2;   // Line 2
    case 2: return 3;   // Triangle 3
    case 3: return 4;   // Quadrangle 4
    case 15: return 1;  // Point
    default: return 0;
    }
}

void expect(std::istream& is, const char* token, const char* what)
{
    std::string buf;
    is >> buf;
    if (buf != token) throw std::runtime_error(std::string(what) + " (found '" + buf + "')");
}

void push_element(GlobalMesh& G, int nranks, int number, int partition, std::vector<int> const& ghosts, int* idx, bool bamg)
{
    if (bamg) std::next_permutation(idx + 1, idx + 3);            // mesh.ordering=bamg (gmshmesh.cpp:343-346)
    G.number.push_back(number);
    G.part.push_back(((partition % nranks) + nranks) % nranks);
    for (int g : ghosts) G.gval.push_back(((g % nranks) + nranks) % nranks);
    G.gptr.push_back((int)G.gval.size());
    G.tri.push_back(idx[0]); G.tri.push_back(idx[1]); G.tri.push_back(idx[2]);
}

void read_msh(std::string const& path, std::string const& format, bool bamg, int rank, int nranks, GlobalMesh& G)
{
    std::ifstream ifs(path.c_str(), std::ios::in | std::ios::binary);
    if (!ifs.is_open()) throw std::invalid_argument("Invalid file name " + path + " (file not found)");
    bool const binary = (format == "binary");
    if (!binary && format != "ascii") throw std::logic_error("invalid mesh file format");
    std::string buf;
    double version = 2.2;
    bool swap = false;
    ifs >> buf;
    if (buf == "$MeshFormat") {
        std::string theversion;
        int fmt, size;
        ifs >> theversion >> fmt >> size;
        version = std::stod(theversion);
        if (version < 2) throw std::runtime_error("Nextsim supports only Gmsh version >= 2");
        if (binary) {
            if (ifs.get() != '\n') throw std::runtime_error("Invalid character after $MeshFormat line");
            int one = 0;
            ifs.read((char*)&one, sizeof(int));
            if (one != 1) swap = true;
        }
        expect(ifs, "$EndMeshFormat", "invalid file format entry");
        ifs >> buf;
        if (buf == "$PhysicalNames") {
            int nnames;
            ifs >> nnames;
            for (int n = 0; n < nnames; ++n) { int id, topodim; std::string name; ifs >> topodim >> id >> name; }
            expect(ifs, "$EndPhysicalNames", "invalid file format entry");
            ifs >> buf;
        }
    }
    if (!(buf == "$NOD" || buf == "$Nodes" || buf == "$ParametricNodes"))
        throw std::runtime_error("invalid nodes string '" + buf + "' in gmsh importer");
    unsigned int n_nodes = 0;
    ifs >> n_nodes;
    G.nn = (int)n_nodes;
    G.x.assign(n_nodes, 0.); G.y.assign(n_nodes, 0.);
    if (binary) {
        ifs.get();
        struct __attribute__((packed)) Rec { int id; double c[3]; };
        std::vector<Rec> recs(n_nodes);
        ifs.read((char*)recs.data(), (std::streamsize)(n_nodes * sizeof(Rec)));
        if (!ifs) throw std::runtime_error("truncated $Nodes block");
        for (auto& r : recs) {
            if (swap) { swap_bytes((char*)&r.id, sizeof(int), 1); swap_bytes((char*)r.c, sizeof(double), 3); }
            if (r.id < 1 || r.id > (int)n_nodes) throw std::runtime_error("node id out of range");
            G.x[r.id - 1] = r.c[0]; G.y[r.id - 1] = r.c[1];
        }
        ifs.get();
    } else {
        for (unsigned int i = 0; i < n_nodes; ++i) {
            int id; double c0, c1, c2;
            ifs >> id >> c0 >> c1 >> c2;
            if (!ifs || id < 1 || id > (int)n_nodes) throw std::runtime_error("bad node record");
            G.x[id - 1] = c0; G.y[id - 1] = c1;
        }
    }
    expect(ifs, "$EndNodes", "invalid end nodes string");
    expect(ifs, "$Elements", "invalid elements string");
    int numElements = 0;
    ifs >> numElements;
    G.gptr.assign(1, 0);
    int num_edge = 0, num_edge_diff = 0;
    bool first_triangle = true;
    std::vector<int> ghosts;
    auto adjust = [&](int number) {
        if (first_triangle) {
            num_edge_diff = (num_edge == 0) ? 0 : ((number == 1) ? 0 : num_edge);
            first_triangle = false;
        }
        return number - num_edge_diff;
    };
    if (binary) {
        ifs.get();
        int done = 0;
        std::vector<int> data;
        while (done < numElements) {
            int header[3];
            ifs.read((char*)header, 3 * sizeof(int));
            if (!ifs) throw std::runtime_error("truncated $Elements block");
            if (swap) swap_bytes((char*)header, sizeof(int), 3);
            int const type = header[0], numElems = header[1], numTags = header[2];
            int const numVertices = vertices_of_type(type);
            if (numVertices <= 0) throw std::logic_error("Unsupported element type " + std::to_string(type));
            size_t const n = 1 + (size_t)numTags + numVertices;
            if (type != 2) {
                // the reference skips ONE record per header (Gmsh writes one header per non-triangle element,
                // gmshmesh.cpp:577-584); skipping numElems records is the same for such files and also right otherwise
                ifs.seekg((std::streamoff)(sizeof(int) * n * numElems), std::ios::cur);
                done += numElems;
                ++num_edge;
                continue;
            }
            data.resize(n);
            for (int i = 0; i < numElems; ++i) {
                ifs.read((char*)data.data(), (std::streamsize)(sizeof(int) * n));
                if (!ifs) throw std::runtime_error("truncated triangle record");
                if (swap) swap_bytes((char*)data.data(), sizeof(int), n);
                int number = data[0];
                int const numPartitions = (version >= 2.2 && numTags > 3) ? data[3] : 1;
                int const partition = (version < 2.2 && numTags > 2) ? data[3] - 1 : (version >= 2.2 && numTags > 3) ? data[4] - 1 : 0;
                ghosts.clear();
                for (int j = 0; j < numPartitions - 1; ++j) ghosts.push_back((-data[5 + j]) - 1);
                int idx[3] = {data[numTags + 1], data[numTags + 2], data[numTags + 3]};
                push_element(G, nranks, adjust(number), partition, ghosts, idx, bamg);
            }
            done += numElems;
        }
    } else {
        for (int i = 0; i < numElements; ++i) {
            int number, type, numTags;
            ifs >> number >> type >> numTags;
            if (!ifs) throw std::runtime_error("bad element record");
            if (type != 2) {
                ifs.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
                ++num_edge;
                continue;
            }
            int numPartitions = 1, partition = (nranks > 1) ? rank : 0;
            ghosts.clear();
            for (int j = 0; j < numTags; ++j) {
                int tag;
                ifs >> tag;
                if (j == 2 && numTags > 3) numPartitions = tag;
                else if (j == 3) partition = tag - 1;
                else if (j >= 4 && j < 4 + numPartitions - 1) ghosts.push_back((-tag) - 1);
            }
            int idx[3];
            ifs >> idx[0] >> idx[1] >> idx[2];
            push_element(G, nranks, adjust(number), partition, ghosts, idx, bamg);
        }
    }
    expect(ifs, "$EndElements", "invalid end elements string");
    G.ne = (int)G.number.size();
    for (int v : G.tri) if (v < 1 || v > G.nn) throw std::runtime_error("triangle refers to a node outside the file");
}

// ---------------------------------------------------------------------------------------------------
// nodalGrid + initUpdateGhosts for one rank
// ---------------------------------------------------------------------------------------------------
void build_local(GlobalMesh const& G, int me, int nranks, nsx_partmesh& M)
{
    int const nn = G.nn, ne = G.ne;
    M.rank = me; M.nranks = nranks;
    M.global_num_nodes = nn; M.global_num_elements = ne;
    if (nranks == 1) {
        M.num_nodes = M.local_ndof = nn;
        M.num_elements = M.local_nelements = ne;
        M.x = G.x; M.y = G.y; M.indices = G.tri;
        M.ghost_nodes.assign(3 * (size_t)ne, 0);
        M.node_gid.resize(nn); M.node_rid.resize(nn);
        for (int i = 0; i < nn; ++i) M.node_gid[i] = M.node_rid[i] = i + 1;
        M.elem_gid = G.number;
        M.elem_part.assign(ne, 0);
        M.send_ptr.assign(1, 0); M.recv_ptr.assign(1, 0);
        return;
    }
    // elements by partition and by ghost tag, file order
    std::vector<int> own_ptr(nranks + 1, 0), gh_ptr(nranks + 1, 0);
    for (int e = 0; e < ne; ++e) {
        own_ptr[G.part[e] + 1]++;
        for (int q = G.gptr[e]; q < G.gptr[e + 1]; ++q)
            if (G.gval[q] != G.part[e]) {
                bool dup = false;
                for (int q2 = G.gptr[e]; q2 < q; ++q2) dup |= (G.gval[q2] == G.gval[q]);
                if (!dup) gh_ptr[G.gval[q] + 1]++;
            }
    }
    for (int r = 0; r < nranks; ++r) { own_ptr[r + 1] += own_ptr[r]; gh_ptr[r + 1] += gh_ptr[r]; }
    std::vector<int> own_e(own_ptr[nranks]), gh_e(gh_ptr[nranks]);
    {
        std::vector<int> fo(own_ptr.begin(), own_ptr.end() - 1), fg(gh_ptr.begin(), gh_ptr.end() - 1);
        for (int e = 0; e < ne; ++e) {
            own_e[fo[G.part[e]]++] = e;
            for (int q = G.gptr[e]; q < G.gptr[e + 1]; ++q)
                if (G.gval[q] != G.part[e]) {
                    bool dup = false;
                    for (int q2 = G.gptr[e]; q2 < q; ++q2) dup |= (G.gval[q2] == G.gval[q]);
                    if (!dup) gh_e[fg[G.gval[q]]++] = e;
                }
        }
    }
    auto node = [&](int e, int i) { return G.tri[3 * (size_t)e + i] - 1; };

    // pass 1: provisional owned nodes of every rank (gmshmesh.cpp:907-941), owner = lowest rank (1057-1094)
    std::vector<int> stamp_gn(nn, -1), stamp_pv(nn, -1), owner(nn, -1);
    std::vector<int> prov_ptr(nranks + 1, 0), prov;
    prov.reserve((size_t)nn + nn / 8);
    for (int r = 0; r < nranks; ++r) {
        for (int k = gh_ptr[r]; k < gh_ptr[r + 1]; ++k)
            for (int i = 0; i < 3; ++i) stamp_gn[node(gh_e[k], i)] = r;
        for (int k = own_ptr[r]; k < own_ptr[r + 1]; ++k) {
            int const e = own_e[k];
            bool const tagged = G.gptr[e + 1] > G.gptr[e];
            for (int i = 0; i < 3; ++i) {
                int const n = node(e, i);
                if ((stamp_gn[n] != r || tagged) && stamp_pv[n] != r) {
                    stamp_pv[n] = r;
                    prov.push_back(n);
                    if (owner[n] < 0) owner[n] = r;
                }
            }
        }
        prov_ptr[r + 1] = (int)prov.size();
    }
    for (int n = 0; n < nn; ++n)
        if (owner[n] < 0) throw std::invalid_argument("partition tags leave node " + std::to_string(n + 1) + " without an owner");
    // reordered global ids: contiguous per rank, u block then v block (gmshmesh.cpp:1178-1220)
    std::vector<long long> base(nranks + 1, 0);
    for (int n = 0; n < nn; ++n) base[owner[n] + 1]++;
    std::vector<int> n_owned(nranks);
    for (int r = 0; r < nranks; ++r) { n_owned[r] = (int)base[r + 1]; base[r + 1] = base[r] + 2 * base[r + 1]; }
    std::vector<int> rid(nn), own_pos(nn);
    {
        std::vector<int> cnt(nranks, 0);
        for (int n = 0; n < nn; ++n) { own_pos[n] = cnt[owner[n]]++; rid[n] = (int)(base[owner[n]] + 1 + own_pos[n]); }
    }

    // pass 2: ghost set of every rank (for my send lists); the full local mesh for `me`
    std::fill(stamp_pv.begin(), stamp_pv.end(), -1);
    std::vector<int> stamp_loc(nn, -1);
    std::vector<std::vector<int>> send_nodes(nranks);
    std::vector<int> my_ghosts, my_loaded;
    std::vector<int> loaded, locals;
    for (int r = 0; r < nranks; ++r) {
        for (int k = prov_ptr[r]; k < prov_ptr[r + 1]; ++k) stamp_pv[prov[k]] = r;
        // loaded elements in file order = merge of the two ascending lists
        loaded.clear();
        {
            int a = own_ptr[r], b = gh_ptr[r];
            while (a < own_ptr[r + 1] || b < gh_ptr[r + 1]) {
                if (b >= gh_ptr[r + 1] || (a < own_ptr[r + 1] && own_e[a] < gh_e[b])) loaded.push_back(own_e[a++]);
                else loaded.push_back(gh_e[b++]);
            }
        }
        locals.clear();
        for (int e : loaded) {
            if (G.part[e] < r) continue;
            if (stamp_pv[node(e, 0)] != r && stamp_pv[node(e, 1)] != r && stamp_pv[node(e, 2)] != r) continue;
            for (int i = 0; i < 3; ++i) {
                int const n = node(e, i);
                if (stamp_loc[n] != r) { stamp_loc[n] = r; locals.push_back(n); }
            }
        }
        // ghosts of r: local but not provisional, or provisional but owned by a lower rank
        auto ghost_of_r = [&](int n) { return stamp_pv[n] != r || owner[n] != r; };
        if (r == me) {
            for (int n : locals) if (stamp_pv[n] != r) my_ghosts.push_back(n);
            for (int k = prov_ptr[r]; k < prov_ptr[r + 1]; ++k) if (owner[prov[k]] != r) my_ghosts.push_back(prov[k]);
            my_loaded = loaded;
        } else {
            for (int n : locals) if (stamp_pv[n] != r && owner[n] == me) send_nodes[r].push_back(n);
            for (int k = prov_ptr[r]; k < prov_ptr[r + 1]; ++k) if (owner[prov[k]] == me) send_nodes[r].push_back(prov[k]);
        }
        (void)ghost_of_r;
    }
    // my local numbering: owned ascending file id, then ghosts ascending file id (gmshmesh.cpp:1165-1169)
    std::sort(my_ghosts.begin(), my_ghosts.end());
    my_ghosts.erase(std::unique(my_ghosts.begin(), my_ghosts.end()), my_ghosts.end());
    std::vector<int> g2l(nn, -1);
    M.local_ndof = n_owned[me];
    M.num_nodes = M.local_ndof + (int)my_ghosts.size();
    M.node_gid.resize(M.num_nodes);
    for (int n = 0; n < nn; ++n) if (owner[n] == me) { g2l[n] = own_pos[n]; M.node_gid[own_pos[n]] = n + 1; }
    for (size_t j = 0; j < my_ghosts.size(); ++j) { g2l[my_ghosts[j]] = M.local_ndof + (int)j; M.node_gid[M.local_ndof + j] = my_ghosts[j] + 1; }
    M.x.resize(M.num_nodes); M.y.resize(M.num_nodes); M.node_rid.resize(M.num_nodes);
    for (int l = 0; l < M.num_nodes; ++l) {
        int const n = M.node_gid[l] - 1;
        M.x[l] = G.x[n]; M.y[l] = G.y[n]; M.node_rid[l] = rid[n];
    }
    // kept elements (gmshmesh.cpp:1271-1312, 1384-1417)
    std::vector<int> kept_own, kept_gh;
    for (int e : my_loaded) {
        if (G.part[e] < me) continue;
        if (g2l[node(e, 0)] < 0 || g2l[node(e, 1)] < 0 || g2l[node(e, 2)] < 0) continue;
        (G.part[e] == me ? kept_own : kept_gh).push_back(e);
    }
    M.local_nelements = (int)kept_own.size();
    M.num_elements = (int)(kept_own.size() + kept_gh.size());
    M.indices.resize(3 * (size_t)M.num_elements); M.ghost_nodes.resize(3 * (size_t)M.num_elements);
    M.elem_gid.resize(M.num_elements); M.elem_part.resize(M.num_elements);
    {
        size_t c = 0;
        for (auto const* v : {&kept_own, &kept_gh})
            for (int e : *v) {
                for (int i = 0; i < 3; ++i) {
                    int const l = g2l[node(e, i)] + 1;
                    M.indices[3 * c + i] = l;
                    M.ghost_nodes[3 * c + i] = (unsigned char)(l > M.local_ndof);
                }
                M.elem_gid[c] = G.number[e]; M.elem_part[c] = G.part[e];
                ++c;
            }
    }
    // M_local_ghosts_local_index: my ghosts by ascending reordered id, grouped by owner
    std::vector<int> gs = my_ghosts;
    std::sort(gs.begin(), gs.end(), [&](int a, int b) { return rid[a] < rid[b]; });
    M.local_ghost.resize(gs.size());
    M.recv_ptr.assign(1, 0);
    for (size_t j = 0; j < gs.size(); ++j) {
        M.local_ghost[j] = rid[gs[j]];
        int const o = owner[gs[j]];
        if (M.recv_peer.empty() || M.recv_peer.back() != o) { if (!M.recv_peer.empty()) M.recv_ptr.push_back((int)j); M.recv_peer.push_back(o); }
        M.recv_idx.push_back(g2l[gs[j]]);
    }
    if (!M.recv_peer.empty()) M.recv_ptr.push_back((int)gs.size());
    // M_extract_local_index: my owned nodes that rank q holds as ghosts, ascending reordered (= file) id
    M.send_ptr.assign(1, 0);
    for (int q = 0; q < nranks; ++q) {
        auto& v = send_nodes[q];
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        M.send_peer.push_back(q);
        for (int n : v) M.send_idx.push_back(own_pos[n]);
        M.send_ptr.push_back((int)M.send_idx.size());
    }
}

// ---------------------------------------------------------------------------------------------------
// bamg's NodalElementConnectivity / NodalConnectivity in bamg's chain order
// ---------------------------------------------------------------------------------------------------
void bamg_tables(nsx_partmesh& M)
{
    int const nn = M.num_nodes, ne = M.num_elements;
    double const NaN = std::numeric_limits<double>::quiet_NaN();
    // node -> element: head-insertion chains give DESCENDING element ids (Mesh.cpp:526-537, 804-811)
    std::vector<int> deg(nn, 0);
    for (size_t k = 0; k < 3 * (size_t)ne; ++k) deg[M.indices[k] - 1]++;
    int w = 0;
    for (int d : deg) w = std::max(w, d);
    M.nec_w = w;
    M.nec.assign((size_t)nn * w, NaN);
    {
        std::vector<int> fill(nn, 0);
        for (int e = ne - 1; e >= 0; --e)
            for (int i = 0; i < 3; ++i) {
                int const n = M.indices[3 * (size_t)e + i] - 1;
                M.nec[(size_t)n * w + fill[n]++] = e + 1;
            }
    }
    // edges numbered by first appearance over (triangle, local edge); local edge k joins the vertices
    // VerticesOfTriangularEdge[k] = {1,2},{2,0},{0,1} (contrib/bamg/include/macros.h:13, Mesh.cpp:583-606)
    static const int vote[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    // bucket the 3*ne (triangle, local edge) records by their lower node (counting sort), find in every small bucket
    // the first appearance of each distinct upper node, and number the edges by scanning the appearance slots in
    // order: O(ne) with tiny local sorts instead of one global sort
    struct Half { int hi, order; };
    std::vector<int> bstart(nn + 1, 0);
    for (int e = 0; e < ne; ++e)
        for (int k = 0; k < 3; ++k) {
            int const a = M.indices[3 * (size_t)e + vote[k][0]] - 1, b = M.indices[3 * (size_t)e + vote[k][1]] - 1;
            bstart[std::min(a, b) + 1]++;
        }
    for (int n = 0; n < nn; ++n) bstart[n + 1] += bstart[n];
    std::vector<Half> half(3 * (size_t)ne);
    {
        std::vector<int> fill(bstart.begin(), bstart.end() - 1);
        for (int e = 0; e < ne; ++e)
            for (int k = 0; k < 3; ++k) {
                int const a = M.indices[3 * (size_t)e + vote[k][0]] - 1, b = M.indices[3 * (size_t)e + vote[k][1]] - 1;
                half[fill[std::min(a, b)]++] = {std::max(a, b), 3 * e + k};
            }
    }
    std::vector<int> first_lo(3 * (size_t)ne, -1), first_hi(3 * (size_t)ne, -1);     // appearance slot -> new edge
    for (int n = 0; n < nn; ++n) {
        Half* b = half.data() + bstart[n];
        int const len = bstart[n + 1] - bstart[n];
        std::sort(b, b + len, [](Half const& p, Half const& q) { return p.hi != q.hi ? p.hi < q.hi : p.order < q.order; });
        for (int j = 0; j < len; ++j)
            if (j == 0 || b[j].hi != b[j - 1].hi) { first_lo[b[j].order] = n; first_hi[b[j].order] = b[j].hi; }
    }
    struct Rec { unsigned long long key; int order; };
    std::vector<Rec> uniq;
    uniq.reserve(3 * (size_t)ne / 2 + 16);
    for (size_t o = 0; o < 3 * (size_t)ne; ++o)
        if (first_lo[o] >= 0) uniq.push_back({(unsigned long long)first_lo[o] * (unsigned long long)nn + (unsigned long long)first_hi[o], (int)o});
    { std::vector<Half>().swap(half); std::vector<int>().swap(first_lo); std::vector<int>().swap(first_hi); }
    // node -> node: chains again, DESCENDING edge id (Mesh.cpp:830-865); last column = count
    std::vector<int> deg2(nn, 0);
    for (auto const& u : uniq) { deg2[(int)(u.key / nn)]++; deg2[(int)(u.key % nn)]++; }
    int w2 = 0;
    for (int d : deg2) w2 = std::max(w2, d);
    w2 += 1;
    M.nc_w = w2;
    M.nc.assign((size_t)nn * w2, 0.);
    {
        std::vector<int> fill(nn, 0);
        for (size_t j = uniq.size(); j-- > 0;) {
            int const lo = (int)(uniq[j].key / nn), hi = (int)(uniq[j].key % nn);
            M.nc[(size_t)lo * w2 + fill[lo]++] = hi + 1;
            M.nc[(size_t)hi * w2 + fill[hi]++] = lo + 1;
        }
        for (int n = 0; n < nn; ++n) M.nc[(size_t)n * w2 + w2 - 1] = deg2[n];
    }
}

nsx_partmesh* finish(GlobalMesh const& G, int rank, int nranks)
{
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("rank / nranks out of range");
    auto* M = new nsx_partmesh();
    try {
        build_local(G, rank, nranks, *M);
        bamg_tables(*M);
        M->lat.assign(M->num_nodes, 0.);
        M->mask_dirichlet.assign(M->num_nodes, 0);
    } catch (...) { delete M; throw; }
    return M;
}

} // namespace

#define PM_TRY try {
#define PM_CATCH } catch (std::exception const& e) { g_pm_err = e.what(); return 2; } return 0;

extern "C" const char* nsx_partmesh_last_error(void) { return g_pm_err.c_str(); }

extern "C" int nsx_partmesh_read(const char* path, const char* format, const char* ordering, int rank, int nranks,
                                 nsx_partmesh_handle* out)
{
    PM_TRY
    if (!path || !format || !ordering || !out) throw std::invalid_argument("nsx_partmesh_read: NULL argument");
    std::string const ord(ordering);
    if (ord != "gmsh" && ord != "bamg") throw std::invalid_argument("mesh.ordering must be gmsh or bamg");
    GlobalMesh G;
    read_msh(path, format, ord == "bamg", rank, nranks, G);
    *out = finish(G, rank, nranks);
    PM_CATCH
}

extern "C" int nsx_partmesh_build(int nn, const double* x, const double* y, int ne, const int* tri, const int* partition,
                                  const int* ghost_ptr, const int* ghost_val, int rank, int nranks, nsx_partmesh_handle* out)
{
    PM_TRY
    if (!x || !y || !tri || !out || nn <= 0 || ne <= 0) throw std::invalid_argument("nsx_partmesh_build: bad argument");
    if (nranks > 1 && (!partition || !ghost_ptr)) throw std::invalid_argument("nsx_partmesh_build: partition tags required");
    GlobalMesh G;
    G.nn = nn; G.ne = ne;
    G.x.assign(x, x + nn); G.y.assign(y, y + nn);
    G.tri.assign(tri, tri + 3 * (size_t)ne);
    for (int v : G.tri) if (v < 1 || v > nn) throw std::invalid_argument("triangle refers to a node outside the mesh");
    G.number.resize(ne);
    for (int e = 0; e < ne; ++e) G.number[e] = e + 1;
    G.part.assign(ne, 0);
    G.gptr.assign(ne + 1, 0);
    if (nranks > 1) {
        for (int e = 0; e < ne; ++e) G.part[e] = ((partition[e] % nranks) + nranks) % nranks;
        G.gptr.assign(ghost_ptr, ghost_ptr + ne + 1);
        G.gval.resize(G.gptr[ne]);
        for (int q = 0; q < G.gptr[ne]; ++q) G.gval[q] = ((ghost_val[q] % nranks) + nranks) % nranks;
    }
    *out = finish(G, rank, nranks);
    PM_CATCH
}

extern "C" int nsx_partmesh_destroy(nsx_partmesh_handle h) { delete h; return 0; }

// FiniteElement::bcMarkedNodes (FE.cpp:150-271): flags are 1-based ROOT node ids; the Dirichlet mask covers owned
// nodes only, the Neumann flags include ghosts; both sorted unique 0-based local ids.
extern "C" int nsx_partmesh_bc_marked_nodes(nsx_partmesh_handle M, const int* dirichlet_flags_root, int n_dirichlet,
                                            const int* neumann_flags_root, int n_neumann)
{
    PM_TRY
    if (!M) throw std::invalid_argument("NULL handle");
    std::vector<int> g2l((size_t)M->global_num_nodes + 1, -1);
    for (int l = 0; l < M->num_nodes; ++l) g2l[M->node_gid[l]] = l;
    auto local = [&](int id) { return (id >= 1 && id <= M->global_num_nodes) ? g2l[id] : -1; };
    M->dirichlet_flags.clear(); M->neumann_flags.clear();
    for (int k = 0; k < n_dirichlet; ++k) { int const l = local(dirichlet_flags_root[k]); if (l >= 0 && l < M->local_ndof) M->dirichlet_flags.push_back(l); }
    for (int k = 0; k < n_neumann; ++k) { int const l = local(neumann_flags_root[k]); if (l >= 0) M->neumann_flags.push_back(l); }
    for (auto* v : {&M->dirichlet_flags, &M->neumann_flags}) { std::sort(v->begin(), v->end()); v->erase(std::unique(v->begin(), v->end()), v->end()); }
    M->mask_dirichlet.assign(M->num_nodes, 0);
    for (int l : M->dirichlet_flags) M->mask_dirichlet[l] = 1;
    PM_CATCH
}

extern "C" int nsx_partmesh_set_lat(nsx_partmesh_handle M, const double* lat_local)
{
    PM_TRY
    if (!M || !lat_local) throw std::invalid_argument("NULL argument");
    M->lat.assign(lat_local, lat_local + M->num_nodes);
    PM_CATCH
}

extern "C" int nsx_partmesh_lat_from_mpp(nsx_partmesh_handle M, const char* mppfile)
{
    PM_TRY
    if (!M || !mppfile) throw std::invalid_argument("NULL argument");
    M->lat.assign(M->num_nodes, 0.);
    if (nsx_mapx_latlon(mppfile, M->num_nodes, M->x.data(), M->y.data(), M->lat.data(), nullptr) != 0)
        throw std::runtime_error(nsx_mapx_last_error());
    PM_CATCH
}

// Fills the structs nsx_create() takes; the pointers stay valid until nsx_partmesh_destroy.
extern "C" int nsx_partmesh_views(nsx_partmesh_handle M, NsxMesh* mesh, NsxHalo* halo)
{
    PM_TRY
    if (!M || !mesh) throw std::invalid_argument("NULL argument");
    mesh->num_nodes = M->num_nodes; mesh->local_ndof = M->local_ndof;
    mesh->num_elements = M->num_elements; mesh->local_nelements = M->local_nelements;
    mesh->coord_x = M->x.data(); mesh->coord_y = M->y.data(); mesh->indices = M->indices.data();
    mesh->ghost_nodes = M->ghost_nodes.data(); mesh->mask_dirichlet = M->mask_dirichlet.data();
    mesh->neumann_flags = M->neumann_flags.data(); mesh->n_neumann_flags = (int)M->neumann_flags.size();
    mesh->nodal_element_connectivity = M->nec.data(); mesh->nec_width = M->nec_w;
    mesh->nodal_connectivity = M->nc.data(); mesh->nc_width = M->nc_w;
    mesh->lat = M->lat.data();
    if (halo) {
        halo->rank = M->rank; halo->nranks = M->nranks;
        halo->n_send_peers = (int)M->send_peer.size(); halo->send_peer = M->send_peer.data();
        halo->send_ptr = M->send_ptr.data(); halo->send_idx = M->send_idx.data();
        halo->n_recv_peers = (int)M->recv_peer.size(); halo->recv_peer = M->recv_peer.data();
        halo->recv_ptr = M->recv_ptr.data(); halo->recv_idx = M->recv_idx.data();
    }
    PM_CATCH
}

// ids[0..3] = local -> file node id (1-based), local -> reordered global id, local element -> file element number,
// local element -> partition; sizes[0..3] = global nodes, global triangles, ghost nodes, dirichlet flags
extern "C" int nsx_partmesh_ids(nsx_partmesh_handle M, const int** ids, int* sizes)
{
    PM_TRY
    if (!M || !ids || !sizes) throw std::invalid_argument("NULL argument");
    ids[0] = M->node_gid.data(); ids[1] = M->node_rid.data(); ids[2] = M->elem_gid.data(); ids[3] = M->elem_part.data();
    sizes[0] = M->global_num_nodes; sizes[1] = M->global_num_elements;
    sizes[2] = M->num_nodes - M->local_ndof; sizes[3] = (int)M->dirichlet_flags.size();
    PM_CATCH
}
