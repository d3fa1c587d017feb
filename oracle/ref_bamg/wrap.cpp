// oracle/ref_bamg/wrap.cpp -- TEST INFRASTRUCTURE ONLY.
// Thin C entry point around the UNMODIFIED reference bamg library (contrib/bamg, compiled from the sources where
// they lie under /root/reference by oracle/ref_bamg/Makefile into oracle/_ref/libref_bamg.so).  It makes the call
// FiniteElement::distributedMeshProcessing makes (model/finiteelement.cpp:77-80):
//     BamgConvertMeshx(bamgmesh, bamggeom, &indexTr[0], &coordX[0], &coordY[0], numNodes, numTriangles)
// and hands back the two tables the hot path reads (bamgmesh->NodalConnectivity, ->NodalElementConnectivity),
// so that the oracle's restatement of bamg's chain orders (and the product's host library) are pinned against
// the reference's own code.
#include <cstring>
#include "BamgConvertMeshx.h"
#include "BamgMesh.h"
#include "BamgGeom.h"

extern "C" {

struct RefBamgTables { BamgMesh* mesh; BamgGeom* geom; };

void* ref_bamg_convert(const int* index, const double* x, const double* y, int nods, int nels, int* sizes)
{
    RefBamgTables* t = new RefBamgTables();
    t->mesh = new BamgMesh();
    t->geom = new BamgGeom();
    BamgConvertMeshx(t->mesh, t->geom, const_cast<int*>(index), const_cast<double*>(x), const_cast<double*>(y), nods, nels);
    sizes[0] = t->mesh->NodalConnectivitySize[0];        sizes[1] = t->mesh->NodalConnectivitySize[1];
    sizes[2] = t->mesh->NodalElementConnectivitySize[0]; sizes[3] = t->mesh->NodalElementConnectivitySize[1];
    sizes[4] = t->mesh->ElementConnectivitySize[0];      sizes[5] = t->mesh->ElementConnectivitySize[1];
    return t;
}

void ref_bamg_get(void* h, double* nodal_connectivity, double* nodal_element_connectivity, double* element_connectivity)
{
    RefBamgTables* t = (RefBamgTables*)h;
    BamgMesh* m = t->mesh;
    if (nodal_connectivity)
        std::memcpy(nodal_connectivity, m->NodalConnectivity, sizeof(double) * m->NodalConnectivitySize[0] * m->NodalConnectivitySize[1]);
    if (nodal_element_connectivity)
        std::memcpy(nodal_element_connectivity, m->NodalElementConnectivity,
                    sizeof(double) * m->NodalElementConnectivitySize[0] * m->NodalElementConnectivitySize[1]);
    if (element_connectivity)
        std::memcpy(element_connectivity, m->ElementConnectivity, sizeof(double) * m->ElementConnectivitySize[0] * m->ElementConnectivitySize[1]);
}

void ref_bamg_free(void* h)
{
    RefBamgTables* t = (RefBamgTables*)h;
    delete t->mesh;
    delete t->geom;
    delete t;
}

}
