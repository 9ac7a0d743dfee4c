/*
 * frt_leafruns.h -- bounding groups over the long triangle runs of the reference's divided tree (host side, scene upload).
 *
 * group_divide (reference src/shapes/group.c:300-370) turns a group into [left half, right half, every child that
 * straddles the split plane ...].  The straddlers stay direct children, and the reference gives a triangle no bounding
 * box of its own (triangle.c:11-45 is called for every child of a group whose box the ray crosses): the six dragons of
 * bounding_boxes.yml carry 200 .. 270 straddling triangles in each of the groups of their upper levels, so a ray that
 * enters a dragon's box pays thousands of Moeller-Trumbore tests before it reaches a small group.
 *
 * The flattened tree is a PRE-ORDER list whose group nodes do nothing but cull; a leaf is looked at in the reference's
 * order no matter how many culling nodes sit in front of it.  So the upload inserts groups of its own over CONSECUTIVE
 * triangle children (the order of the leaves, which shadow rays depend on -- group.c:105-123 -- is untouched): a binary
 * tree per run, split where the surface-area cost of the two ordered halves is smallest, down to FRT_RUN_LEAF triangles.
 * A culled group drops only triangles whose own vertices lie inside its box, i.e. triangles the ray does not touch;
 * the boxes are padded far beyond the FP64 rounding of the triangle test.  Normals, materials and patterns use the
 * composite transform of a leaf, never the parent chain, so the inserted groups are invisible to shading.
 */
#ifndef FRT_LEAFRUNS_H
#define FRT_LEAFRUNS_H

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "../../include/frt_b200.h"

#define FRT_RUN_MIN 6  /* shortest run of consecutive triangle children that gets groups of its own */
#define FRT_RUN_LEAF 4 /* triangles left side by side under the lowest inserted group */

namespace frt_leafruns {

struct Box {
    double lo[3], hi[3];
    void clear()
    {
        for (int k = 0; k < 3; ++k) {
            lo[k] = INFINITY;
            hi[k] = -INFINITY;
        }
    }
    void add(const Box &b)
    {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
    double area() const
    {
        const double x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
        return x * y + y * z + z * x;
    }
};

/* bounds of a triangle's three vertices in its own space (prim_params: p1 p2 p3 ...), padded: the pad is nine orders
 * above the rounding of the FP64 triangle test and far below anything a picture can show */
inline Box
triangle_box(const double *prm)
{
    Box b;
    double m = 1.0;
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = std::min(prm[k], std::min(prm[3 + k], prm[6 + k]));
        b.hi[k] = std::max(prm[k], std::max(prm[3 + k], prm[6 + k]));
        m = std::max(m, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
    }
    const double pad = 1e-9 * m;
    for (int k = 0; k < 3; ++k) {
        b.lo[k] -= pad;
        b.hi[k] += pad;
    }
    return b;
}

struct Builder {
    const frt_scene_desc *d;
    std::vector<frt_node> out;
    std::vector<int32_t> map; /* old index -> new index */
    std::vector<Box> boxes;   /* scratch: boxes of the run at hand */
    std::vector<Box> suffix;
    long inserted = 0;
    int run_min = FRT_RUN_MIN, run_leaf = FRT_RUN_LEAF;
    bool bad = false; /* the description is not the tree it claims to be: upload it as it is */

    static bool is_tri(const frt_node &n) { return n.type == FRT_TRIANGLE || n.type == FRT_SMOOTH_TRIANGLE; }

    void copy_node(int i, int parent_new)
    {
        map[i] = (int32_t)out.size();
        frt_node n = d->nodes[i];
        n.parent = parent_new;
        out.push_back(n);
    }

    /* the triangles run[a .. b) (old indices of consecutive children; boxes[j] belongs to run[j]) */
    void emit_run(const std::vector<int> &run, int a, int b, int parent_new, bool wrap)
    {
        int g = parent_new;
        if (wrap) {
            Box all;
            all.clear();
            for (int j = a; j < b; ++j) {
                all.add(boxes[j]);
            }
            frt_node n{};
            n.type = FRT_GROUP;
            n.skip = 0;
            n.parent = parent_new;
            n.xform = d->nodes[run[a]].xform;
            n.material = -1;
            n.param = -1;
            n.csg_op = 0;
            n.right = 0;
            for (int k = 0; k < 3; ++k) {
                n.bbox_min[k] = all.lo[k];
                n.bbox_max[k] = all.hi[k];
            }
            g = (int)out.size();
            out.push_back(n);
            ++inserted;
        }
        if (b - a <= run_leaf) {
            for (int j = a; j < b; ++j) {
                copy_node(run[j], g);
                out.back().skip = (int32_t)out.size();
            }
        } else {
            /* ordered split with the smallest surface-area cost */
            suffix.resize((size_t)(b - a) + 1);
            Box acc;
            acc.clear();
            for (int j = b - 1; j >= a; --j) {
                acc.add(boxes[j]);
                suffix[j - a] = acc;
            }
            acc.clear();
            int best = (a + b) / 2;
            double best_cost = INFINITY;
            for (int k = a + 1; k < b; ++k) {
                acc.add(boxes[k - 1]);
                const double cost = acc.area() * (k - a) + suffix[k - a].area() * (b - k);
                if (cost < best_cost) {
                    best_cost = cost;
                    best = k;
                }
            }
            emit_run(run, a, best, g, true);
            emit_run(run, best, b, g, true);
        }
        if (wrap) {
            out[g].skip = (int32_t)out.size();
        }
    }

    void emit(int i, int parent_new, bool under_csg)
    {
        const frt_node &n = d->nodes[i];
        const int me = (int)out.size();
        copy_node(i, parent_new);
        if (n.type < FRT_CSG) {
            out[me].skip = me + 1;
            return;
        }
        std::vector<int> kids;
        for (int c = i + 1; c < n.skip; c = d->nodes[c].skip) {
            kids.push_back(c);
        }
        const bool plain = n.type == FRT_GROUP && !under_csg;
        size_t k = 0;
        while (k < kids.size()) {
            size_t e = k;
            if (plain && is_tri(d->nodes[kids[k]])) {
                while (e < kids.size() && is_tri(d->nodes[kids[e]]) && d->nodes[kids[e]].xform == d->nodes[kids[k]].xform) {
                    ++e;
                }
            }
            if (e - k >= (size_t)run_min) {
                std::vector<int> run(kids.begin() + (long)k, kids.begin() + (long)e);
                boxes.resize(run.size());
                for (size_t j = 0; j < run.size(); ++j) {
                    boxes[j] = triangle_box(d->prim_params + d->nodes[run[j]].param);
                }
                /* a run that is the whole child list sits in its parent's box already */
                emit_run(run, 0, (int)run.size(), me, e - k != kids.size());
                k = e;
            } else {
                emit(kids[k], me, under_csg || n.type == FRT_CSG);
                ++k;
            }
        }
        out[me].skip = (int32_t)out.size();
        if (n.type == FRT_CSG) {
            if (n.right < 0 || n.right >= d->n_nodes || map[n.right] < 0) {
                bad = true;
            } else {
                out[me].right = map[n.right];
            }
        }
    }
};

/* true when groups were inserted: `nodes` / `roots` then hold the tree to upload instead of d->nodes / d->roots */
inline bool
augment(const frt_scene_desc *d, std::vector<frt_node> &nodes, std::vector<int32_t> &roots, long *inserted)
{
    /* roots are subtrees laid one after the other; anything else (validate_desc lets it through) is left alone */
    int expect = 0;
    for (int r = 0; r < d->n_roots; ++r) {
        if (d->roots[r] != expect) {
            return false;
        }
        expect = d->nodes[expect].skip;
    }
    if (expect != d->n_nodes) {
        return false;
    }
    Builder b;
    b.d = d;
    if (const char *v = getenv("FRT_RUN_MIN")) { /* development aid: the two thresholds at run time */
        b.run_min = std::max(2, atoi(v));
    }
    if (const char *v = getenv("FRT_RUN_LEAF")) {
        b.run_leaf = std::max(1, atoi(v));
    }
    b.map.assign((size_t)d->n_nodes, -1);
    b.out.reserve((size_t)d->n_nodes + (size_t)d->n_nodes / 2);
    roots.clear();
    for (int r = 0; r < d->n_roots; ++r) {
        roots.push_back((int32_t)b.out.size());
        b.emit(d->roots[r], -1, false);
    }
    if (b.inserted == 0 || b.bad) {
        return false;
    }
    nodes.swap(b.out);
    if (inserted != nullptr) {
        *inserted = b.inserted;
    }
    return true;
}

} /* namespace frt_leafruns */

#endif
