/*
 * mpi.h -- an in-process stand-in for the ~30 MPI symbols GridTools' gcl uses (option ii).
 *
 * TEST INFRASTRUCTURE ONLY.  The build container and the GPU box have no MPI installation and no network, but
 * the reference's gcl (gcl/GCL.hpp:12, gcl/low_level/proc_grids_3D.hpp, gcl/low_level/Halo_Exchange_3D.hpp)
 * includes <mpi.h>.  With this header on the include path the UNMODIFIED reference gcl compiles and runs with
 * THREADS AS RANKS: mpi_shim::run(n, body) starts n threads, thread r is rank r of MPI_COMM_WORLD.
 *
 *   - point-to-point: MPI_Isend copies the message eagerly into a mailbox keyed (communicator context, source,
 *     destination, tag); MPI_Irecv only records the request; MPI_Wait on a receive blocks until the matching
 *     message is in the mailbox.  Messages with the same key are matched in posting order, as MPI guarantees.
 *   - Cartesian topology: row-major ranks like every MPI implementation, rank = (c0*d1 + c1)*d2 + c2, no reordering.
 *   - MPI_Allgather / MPI_Barrier / MPI_Bcast: a generation-counted rendezvous of all ranks of the communicator.
 *   - derived datatypes: only their size is tracked (the halo-exchange path sends MPI_CHAR buffers; the
 *     subarray types empty_field_base.hpp builds are never used for transport there).
 *
 * Header-only (C++17 inline variables) so that both oracle/_ref/libgtref.so and the test executables under
 * tests/_build/ can use it without another library.
 */
#ifndef GTB200_ORACLE_MPI_SHIM_H
#define GTB200_ORACLE_MPI_SHIM_H

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <tuple>
#include <vector>

#define MPI_SUCCESS 0
#define MPI_ORDER_C 56
#define MPI_ORDER_FORTRAN 57
#define MPI_MAX_PROCESSOR_NAME 64
#define MPI_ANY_TAG (-1)
#define MPI_PROC_NULL (-2)
#define MPI_IN_PLACE ((void *)1)

typedef int MPI_Datatype; /* index into mpi_shim::state().type_size */
#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5
#define MPI_UNSIGNED 6
#define MPI_LONG 7
#define MPI_UNSIGNED_LONG 8
#define MPI_LONG_LONG 9

typedef int MPI_Op;
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3

namespace mpi_shim {
    struct comm_t {
        int context = 0; /* 0 = world; dup / cart_create give every rank the same new context */
        int size = 1;
        bool cart = false;
        int ndims = 0;
        int dims[3] = {1, 1, 1};
        int periods[3] = {0, 0, 0};
    };
} // namespace mpi_shim

typedef mpi_shim::comm_t *MPI_Comm;
#define MPI_COMM_NULL ((MPI_Comm)0)
#define MPI_COMM_WORLD (mpi_shim::world())

struct MPI_Status {
    int MPI_SOURCE, MPI_TAG, MPI_ERROR;
};
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)

namespace mpi_shim {
    struct request_t {
        bool is_recv = false;
        void *buf = nullptr;
        size_t bytes = 0;
        int src = 0, dst = 0, tag = 0, context = 0;
    };
} // namespace mpi_shim
typedef mpi_shim::request_t *MPI_Request;
#define MPI_REQUEST_NULL ((MPI_Request)0)

namespace mpi_shim {
    using key_t = std::tuple<int, int, int, int>; /* context, src, dst, tag */

    struct state_t {
        std::mutex m;
        std::condition_variable cv;
        std::map<key_t, std::deque<std::vector<char>>> mailbox;
        std::vector<size_t> type_size{0, 1, 1, sizeof(int), sizeof(float), sizeof(double), sizeof(unsigned),
            sizeof(long), sizeof(unsigned long), sizeof(long long)};
        int world_size = 1;
        bool initialized = false;
        /* collectives: one rendezvous area per communicator context */
        struct coll_t {
            int arrived = 0;
            long generation = 0;
            std::vector<char> data;
        };
        std::map<int, coll_t> coll;
        /* context ids handed out collectively: the n-th dup/cart_create of a parent context gets the same id on
         * every rank */
        std::map<std::pair<int, int>, int> child_context;
        int next_context = 1;
        comm_t world_comm;
    };

    inline state_t &state() {
        static state_t s;
        return s;
    }

    inline int &rank_ref() {
        static thread_local int r = 0;
        return r;
    }
    inline std::map<int, int> &dup_counter() { /* per rank: how many children each context already has */
        static thread_local std::map<int, int> c;
        return c;
    }

    inline MPI_Comm world() {
        state_t &s = state();
        s.world_comm.size = s.world_size;
        return &s.world_comm;
    }

    inline int comm_size(MPI_Comm comm) { return comm->context == 0 ? state().world_size : comm->size; }

    inline int new_context(int parent) {
        state_t &s = state();
        int nth = dup_counter()[parent]++;
        std::lock_guard<std::mutex> l(s.m);
        auto it = s.child_context.find({parent, nth});
        if (it != s.child_context.end())
            return it->second;
        int c = s.next_context++;
        s.child_context[{parent, nth}] = c;
        return c;
    }

    /* all ranks of the communicator meet here; rank 0..size-1 each deposit `bytes` at offset rank*bytes. */
    inline void rendezvous(MPI_Comm comm, const void *mine, size_t bytes, void *all) {
        state_t &s = state();
        const int n_ranks = comm_size(comm);
        std::unique_lock<std::mutex> l(s.m);
        state_t::coll_t &c = s.coll[comm->context];
        /* a rank may arrive for generation g+1 while slow ranks still copy out generation g: wait for the
         * area to drain (arrived counts down after completion) */
        long gen = c.generation;
        if (c.data.size() < bytes * n_ranks)
            c.data.resize(bytes * n_ranks);
        if (bytes)
            std::memcpy(c.data.data() + size_t(rank_ref()) * bytes, mine, bytes);
        if (++c.arrived == n_ranks) {
            c.arrived = 0;
            ++c.generation;
            if (all && bytes)
                std::memcpy(all, c.data.data(), bytes * n_ranks);
            /* the last arriver must not let the next collective overwrite data before everybody copied it out:
             * the others copy while holding the lock right after wake-up, and a new collective needs the lock
             * and all `size` ranks again, so the area is stable until every rank left */
            s.cv.notify_all();
            return;
        }
        s.cv.wait(l, [&] { return c.generation != gen; });
        if (all && bytes)
            std::memcpy(all, c.data.data(), bytes * n_ranks);
    }

    /* Runs body(rank) on n threads, thread r being rank r.  Exceptions are re-thrown on the caller. */
    template <class F>
    void run(int n, F body) {
        state_t &s = state();
        {
            std::lock_guard<std::mutex> l(s.m);
            s.world_size = n;
            s.mailbox.clear();
            s.coll.clear();
        }
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> err(n);
        for (int r = 0; r < n; ++r)
            th.emplace_back([&, r] {
                rank_ref() = r;
                dup_counter().clear();
                try {
                    body(r);
                } catch (...) {
                    err[r] = std::current_exception();
                    std::fprintf(stderr, "mpi_shim: rank %d threw; aborting (other ranks may be blocked)\n", r);
                    std::abort();
                }
            });
        for (auto &t : th)
            t.join();
        s.world_size = 1;
    }

    /* rendezvous data is copied out by waiters under the lock after generation changed, but a fast rank could
     * start the next collective and overwrite c.data before a slow waiter woke up.  Guard: collectives end with a
     * second phase.  (Kept simple: every public collective below calls rendezvous twice -- data, then drain.) */
    inline void collective(MPI_Comm comm, const void *mine, size_t bytes, void *all) {
        rendezvous(comm, mine, bytes, all);
        rendezvous(comm, nullptr, 0, nullptr);
    }
} // namespace mpi_shim

/* ---------------------------------------------------------------------------------- environment */
inline int MPI_Init(int *, char ***) {
    mpi_shim::state().initialized = true;
    return MPI_SUCCESS;
}
inline int MPI_Initialized(int *flag) {
    *flag = mpi_shim::state().initialized ? 1 : 0;
    return MPI_SUCCESS;
}
inline int MPI_Finalize() { return MPI_SUCCESS; }
inline int MPI_Abort(MPI_Comm, int code) { std::abort(); return code; }
inline double MPI_Wtime() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int MPI_Comm_rank(MPI_Comm, int *rank) {
    *rank = mpi_shim::rank_ref();
    return MPI_SUCCESS;
}
inline int MPI_Comm_size(MPI_Comm comm, int *size) {
    *size = comm->context == 0 ? mpi_shim::state().world_size : comm->size;
    return MPI_SUCCESS;
}
inline int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out) {
    auto *c = new mpi_shim::comm_t(*comm);
    c->size = comm->context == 0 ? mpi_shim::state().world_size : comm->size;
    c->context = mpi_shim::new_context(comm->context);
    *out = c;
    return MPI_SUCCESS;
}
inline int MPI_Comm_free(MPI_Comm *comm) {
    if (*comm && (*comm)->context != 0)
        delete *comm;
    *comm = MPI_COMM_NULL;
    return MPI_SUCCESS;
}

/* ---------------------------------------------------------------------------------- Cartesian topology */
inline int MPI_Dims_create(int nnodes, int ndims, int *dims) {
    int fixed = 1, free_dims = 0;
    for (int d = 0; d < ndims; ++d)
        if (dims[d] > 0)
            fixed *= dims[d];
        else
            ++free_dims;
    if (fixed <= 0 || nnodes % fixed)
        return 1;
    int rest = nnodes / fixed;
    std::vector<int> f(free_dims, 1);
    /* balanced factorisation, non-increasing: peel prime factors largest first onto the smallest entry */
    std::vector<int> primes;
    for (int p = 2; rest > 1; ++p)
        while (rest % p == 0) {
            primes.push_back(p);
            rest /= p;
        }
    std::sort(primes.rbegin(), primes.rend());
    for (int p : primes)
        *std::min_element(f.begin(), f.end()) *= p;
    std::sort(f.rbegin(), f.rend());
    int n = 0;
    for (int d = 0; d < ndims; ++d)
        if (dims[d] <= 0)
            dims[d] = f[n++];
    return MPI_SUCCESS;
}
inline int MPI_Cart_create(MPI_Comm comm, int ndims, const int *dims, const int *periods, int, MPI_Comm *out) {
    auto *c = new mpi_shim::comm_t;
    c->context = mpi_shim::new_context(comm->context);
    c->cart = true;
    c->ndims = ndims;
    c->size = 1;
    for (int d = 0; d < ndims && d < 3; ++d) {
        c->dims[d] = dims[d];
        c->periods[d] = periods[d];
        c->size *= dims[d];
    }
    int have = comm->context == 0 ? mpi_shim::state().world_size : comm->size;
    if (c->size > have) {
        delete c;
        *out = MPI_COMM_NULL;
        return 1;
    }
    *out = c;
    return MPI_SUCCESS;
}
inline int MPI_Cart_coords(MPI_Comm comm, int rank, int maxdims, int *coords) {
    for (int d = comm->ndims - 1; d >= 0; --d) {
        if (d < maxdims)
            coords[d] = rank % comm->dims[d];
        rank /= comm->dims[d];
    }
    return MPI_SUCCESS;
}
inline int MPI_Cart_get(MPI_Comm comm, int maxdims, int *dims, int *periods, int *coords) {
    for (int d = 0; d < maxdims && d < comm->ndims; ++d) {
        dims[d] = comm->dims[d];
        periods[d] = comm->periods[d];
    }
    return MPI_Cart_coords(comm, mpi_shim::rank_ref(), maxdims, coords);
}
inline int MPI_Cart_rank(MPI_Comm comm, const int *coords, int *rank) {
    int r = 0;
    for (int d = 0; d < comm->ndims; ++d) {
        int c = coords[d];
        if (comm->periods[d])
            c = ((c % comm->dims[d]) + comm->dims[d]) % comm->dims[d];
        else if (c < 0 || c >= comm->dims[d])
            return 1;
        r = r * comm->dims[d] + c;
    }
    *rank = r;
    return MPI_SUCCESS;
}

/* ---------------------------------------------------------------------------------- datatypes (size only) */
inline int MPI_Type_size(MPI_Datatype t, int *size) {
    *size = int(mpi_shim::state().type_size[t]);
    return MPI_SUCCESS;
}
inline int MPI_Type_contiguous(int count, MPI_Datatype old, MPI_Datatype *out) {
    auto &s = mpi_shim::state();
    std::lock_guard<std::mutex> l(s.m);
    s.type_size.push_back(s.type_size[old] * count);
    *out = int(s.type_size.size()) - 1;
    return MPI_SUCCESS;
}
inline int MPI_Type_create_subarray(
    int ndims, const int *, const int *subsizes, const int *, int, MPI_Datatype old, MPI_Datatype *out) {
    auto &s = mpi_shim::state();
    std::lock_guard<std::mutex> l(s.m);
    size_t n = s.type_size[old];
    for (int d = 0; d < ndims; ++d)
        n *= subsizes[d];
    s.type_size.push_back(n); /* extent bookkeeping only: never used for transport on the halo-exchange path */
    *out = int(s.type_size.size()) - 1;
    return MPI_SUCCESS;
}
inline int MPI_Type_commit(MPI_Datatype *) { return MPI_SUCCESS; }
inline int MPI_Type_free(MPI_Datatype *t) {
    *t = MPI_DATATYPE_NULL;
    return MPI_SUCCESS;
}

/* ---------------------------------------------------------------------------------- point to point */
inline int MPI_Isend(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request *req) {
    auto &s = mpi_shim::state();
    size_t bytes = size_t(count) * s.type_size[type];
    {
        std::lock_guard<std::mutex> l(s.m);
        const char *p = static_cast<const char *>(buf);
        s.mailbox[{comm->context, mpi_shim::rank_ref(), dest, tag}].emplace_back(p, p + bytes);
    }
    s.cv.notify_all();
    *req = new mpi_shim::request_t; /* already complete */
    return MPI_SUCCESS;
}
inline int MPI_Irecv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Request *req) {
    auto *r = new mpi_shim::request_t;
    r->is_recv = true;
    r->buf = buf;
    r->bytes = size_t(count) * mpi_shim::state().type_size[type];
    r->src = source;
    r->dst = mpi_shim::rank_ref();
    r->tag = tag;
    r->context = comm->context;
    *req = r;
    return MPI_SUCCESS;
}
inline int MPI_Wait(MPI_Request *req, MPI_Status *status) {
    mpi_shim::request_t *r = *req;
    if (!r)
        return MPI_SUCCESS;
    if (r->is_recv) {
        auto &s = mpi_shim::state();
        std::unique_lock<std::mutex> l(s.m);
        mpi_shim::key_t key{r->context, r->src, r->dst, r->tag};
        s.cv.wait(l, [&] {
            auto it = s.mailbox.find(key);
            return it != s.mailbox.end() && !it->second.empty();
        });
        auto &q = s.mailbox[key];
        if (q.front().size() > r->bytes) {
            std::fprintf(stderr, "mpi_shim: message of %zu bytes truncated by a %zu byte receive (src %d dst %d tag %d)\n",
                q.front().size(), r->bytes, r->src, r->dst, r->tag);
            std::abort();
        }
        std::memcpy(r->buf, q.front().data(), q.front().size());
        q.pop_front();
        if (status) {
            status->MPI_SOURCE = r->src;
            status->MPI_TAG = r->tag;
            status->MPI_ERROR = MPI_SUCCESS;
        }
    }
    delete r;
    *req = MPI_REQUEST_NULL;
    return MPI_SUCCESS;
}
inline int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *st) {
    for (int i = 0; i < n; ++i)
        MPI_Wait(reqs + i, st ? st + i : nullptr);
    return MPI_SUCCESS;
}
inline int MPI_Send(const void *buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm) {
    MPI_Request r;
    MPI_Isend(buf, count, type, dest, tag, comm, &r);
    return MPI_Wait(&r, nullptr);
}
inline int MPI_Recv(void *buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status *st) {
    MPI_Request r;
    MPI_Irecv(buf, count, type, source, tag, comm, &r);
    return MPI_Wait(&r, st);
}

/* ---------------------------------------------------------------------------------- collectives */
inline int MPI_Barrier(MPI_Comm comm) {
    mpi_shim::collective(comm, nullptr, 0, nullptr);
    return MPI_SUCCESS;
}
inline int MPI_Allgather(
    const void *send, int scount, MPI_Datatype stype, void *recv, int, MPI_Datatype, MPI_Comm comm) {
    mpi_shim::collective(comm, send, size_t(scount) * mpi_shim::state().type_size[stype], recv);
    return MPI_SUCCESS;
}
inline int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
    size_t bytes = size_t(count) * mpi_shim::state().type_size[type];
    int size;
    MPI_Comm_size(comm, &size);
    std::vector<char> all(bytes * size);
    mpi_shim::collective(comm, buf, bytes, all.data());
    std::memcpy(buf, all.data() + bytes * root, bytes);
    return MPI_SUCCESS;
}
inline int MPI_Allreduce(const void *send, void *recv, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm) {
    size_t esz = mpi_shim::state().type_size[type], bytes = esz * count;
    int size;
    MPI_Comm_size(comm, &size);
    std::vector<char> all(bytes * size);
    mpi_shim::collective(comm, send, bytes, all.data());
    auto reduce = [&](auto *out) {
        using T = std::remove_pointer_t<decltype(out)>;
        const T *in = reinterpret_cast<const T *>(all.data());
        for (int i = 0; i < count; ++i) {
            T v = in[i];
            for (int r = 1; r < size; ++r) {
                T w = in[size_t(r) * count + i];
                v = op == MPI_SUM ? v + w : op == MPI_MAX ? std::max(v, w) : std::min(v, w);
            }
            out[i] = v;
        }
    };
    switch (type) {
    case MPI_INT: reduce(static_cast<int *>(recv)); break;
    case MPI_FLOAT: reduce(static_cast<float *>(recv)); break;
    case MPI_DOUBLE: reduce(static_cast<double *>(recv)); break;
    case MPI_LONG: reduce(static_cast<long *>(recv)); break;
    default: std::fprintf(stderr, "mpi_shim: MPI_Allreduce on datatype %d not supported\n", type); std::abort();
    }
    return MPI_SUCCESS;
}

#endif
