// seq.cu -- recorded call sequences: the host-side time loop of a stencil program as ONE native call.
//
// A GridTools user program drives its time loop from C++ (e.g. tests/regression/gcl/copy_stencil_parallel.cpp:126-145:
// he.pack / he.exchange / he.unpack, then run(spec, backend, grid, fields...)); every one of those calls costs a
// microsecond or two of host time.  A host that reaches this library through ctypes / JNI / cgo pays 3-5 us per call
// instead, which is more than a 256x256x80 stencil leaves (25-60 us per step with three launches and four stream/event
// operations per step).  A gtb_seq records such a loop once -- stencil launches, halo exchanges, event record / wait
// operations between the compute and the communication stream -- and replays any slice of it with one call, at the
// cost the reference's own C++ driver has.  It is a convenience of the boundary, not a scheduler: operations are issued
// in recorded order on the streams they were recorded with.
#include "common.cuh"

#include <vector>

using namespace gtb;

namespace {
    enum op_kind { OP_HD64, OP_HD32, OP_VA64, OP_VA32, OP_TRACERS, OP_HALO, OP_RECORD, OP_WAIT, OP_SGATE, OP_HGATE, OP_MARK, OP_STAMP };

    struct op {
        op_kind kind;
        gtb_field f[5];
        double scalar;
        int ni, nj, nk;
        void *stream;
        gtb_halo *halo;
        std::vector<void *> ptrs;
        std::vector<gtb_field> outs, ins; // prepare_tracers
        int event;
        const void *gate_flag;
        uint64_t gate_value;
        void *gate_post;
    };
} // namespace

struct gtb_seq {
    std::vector<op> ops;
    std::vector<cudaEvent_t> events;
    std::vector<cudaEvent_t> marks; // timing-enabled events (gtb_seq_add_mark)
};

namespace {
    int event_of(gtb_seq *s, int slot, cudaEvent_t *out) {
        if (slot < 0 || slot > 4096)
            return fail(GTB_ERR_ARG, "gtb_seq: event slot %d out of range", slot);
        while ((int)s->events.size() <= slot) {
            cudaEvent_t e;
            GTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            s->events.push_back(e);
        }
        *out = s->events[slot];
        return GTB_OK;
    }
} // namespace

GTB_API int gtb_seq_create(gtb_seq **out) {
    if (!out)
        return fail(GTB_ERR_ARG, "gtb_seq_create: null argument");
    if (!dev())
        return GTB_ERR_CUDA;
    *out = new gtb_seq();
    return GTB_OK;
}

GTB_API int gtb_seq_destroy(gtb_seq *s) {
    if (!s)
        return GTB_OK;
    for (cudaEvent_t e : s->events)
        cudaEventDestroy(e);
    for (cudaEvent_t e : s->marks)
        cudaEventDestroy(e);
    delete s;
    return GTB_OK;
}

GTB_API int gtb_seq_size(const gtb_seq *s) { return s ? (int)s->ops.size() : 0; }

GTB_API int gtb_seq_add_hori_diff(gtb_seq *s, int elem_size, const gtb_field *in, const gtb_field *coeff,
    const gtb_field *out, int ni, int nj, int nk, void *stream) {
    if (!s || !in || !coeff || !out || (elem_size != 4 && elem_size != 8))
        return fail(GTB_ERR_ARG, "gtb_seq_add_hori_diff: bad argument");
    op o{};
    o.kind = elem_size == 8 ? OP_HD64 : OP_HD32;
    o.f[0] = *in, o.f[1] = *coeff, o.f[2] = *out;
    o.ni = ni, o.nj = nj, o.nk = nk;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_vert_adv(gtb_seq *s, int elem_size, const gtb_field *utens_stage, const gtb_field *u_stage,
    const gtb_field *wcon, const gtb_field *u_pos, const gtb_field *utens, double dtr_stage, int ni, int nj, int nk,
    void *stream) {
    if (!s || !utens_stage || !u_stage || !wcon || !u_pos || !utens || (elem_size != 4 && elem_size != 8))
        return fail(GTB_ERR_ARG, "gtb_seq_add_vert_adv: bad argument");
    op o{};
    o.kind = elem_size == 8 ? OP_VA64 : OP_VA32;
    o.f[0] = *utens_stage, o.f[1] = *u_stage, o.f[2] = *wcon, o.f[3] = *u_pos, o.f[4] = *utens;
    o.scalar = dtr_stage;
    o.ni = ni, o.nj = nj, o.nk = nk;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_prepare_tracers(gtb_seq *s, const gtb_field *out, const gtb_field *in, int n_tracers,
    const gtb_field *rho, int ni, int nj, int nk, void *stream) {
    if (!s || !out || !in || !rho || n_tracers < 0)
        return fail(GTB_ERR_ARG, "gtb_seq_add_prepare_tracers: bad argument");
    op o{};
    o.kind = OP_TRACERS;
    o.outs.assign(out, out + n_tracers);
    o.ins.assign(in, in + n_tracers);
    o.f[0] = *rho;
    o.ni = ni, o.nj = nj, o.nk = nk;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_halo_exchange(gtb_seq *s, gtb_halo *h, void *const *fields, int n_fields, void *stream) {
    if (!s || !h || !fields || n_fields < 0)
        return fail(GTB_ERR_ARG, "gtb_seq_add_halo_exchange: bad argument");
    op o{};
    o.kind = OP_HALO;
    o.halo = h;
    o.ptrs.assign(fields, fields + n_fields);
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_stencil_gate(gtb_seq *s, const void *wait_flag, uint64_t wait_value, void *post_counter) {
    if (!s)
        return fail(GTB_ERR_ARG, "gtb_seq_add_stencil_gate: null sequence");
    op o{};
    o.kind = OP_SGATE;
    o.gate_flag = wait_flag, o.gate_value = wait_value, o.gate_post = post_counter;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_halo_gate(gtb_seq *s, gtb_halo *h, const void *counter, uint64_t value) {
    if (!s || !h)
        return fail(GTB_ERR_ARG, "gtb_seq_add_halo_gate: null argument");
    op o{};
    o.kind = OP_HGATE;
    o.halo = h;
    o.gate_flag = counter, o.gate_value = value;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_record(gtb_seq *s, int event, void *stream) {
    if (!s)
        return fail(GTB_ERR_ARG, "gtb_seq_add_record: null sequence");
    cudaEvent_t e;
    int st = event_of(s, event, &e);
    if (st)
        return st;
    op o{};
    o.kind = OP_RECORD;
    o.event = event;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_wait(gtb_seq *s, void *stream, int event) {
    if (!s)
        return fail(GTB_ERR_ARG, "gtb_seq_add_wait: null sequence");
    cudaEvent_t e;
    int st = event_of(s, event, &e);
    if (st)
        return st;
    op o{};
    o.kind = OP_WAIT;
    o.event = event;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_mark(gtb_seq *s, int mark, void *stream) {
    if (!s || mark < 0 || mark > 4096)
        return fail(GTB_ERR_ARG, "gtb_seq_add_mark: bad argument");
    while ((int)s->marks.size() <= mark) {
        cudaEvent_t e;
        GTB_CUDA(cudaEventCreate(&e));
        s->marks.push_back(e);
    }
    op o{};
    o.kind = OP_MARK;
    o.event = mark;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_add_stamp(gtb_seq *s, void *device_u64, void *stream) {
    if (!s || !device_u64)
        return fail(GTB_ERR_ARG, "gtb_seq_add_stamp: bad argument");
    op o{};
    o.kind = OP_STAMP;
    o.gate_post = device_u64;
    o.stream = stream;
    s->ops.push_back(o);
    return GTB_OK;
}

GTB_API int gtb_seq_elapsed_ms(gtb_seq *s, int mark_a, int mark_b, float *ms) {
    if (!s || !ms || mark_a < 0 || mark_b < 0 || mark_a >= (int)s->marks.size() || mark_b >= (int)s->marks.size())
        return fail(GTB_ERR_ARG, "gtb_seq_elapsed_ms: bad argument");
    GTB_CUDA(cudaEventSynchronize(s->marks[mark_b]));
    GTB_CUDA(cudaEventElapsedTime(ms, s->marks[mark_a], s->marks[mark_b]));
    return GTB_OK;
}

GTB_API int gtb_seq_run(gtb_seq *s, int first, int count) {
    if (!s || first < 0 || count < 0 || (size_t)first + (size_t)count > s->ops.size())
        return fail(GTB_ERR_ARG, "gtb_seq_run: slice out of range");
    for (int i = first; i < first + count; ++i) {
        const op &o = s->ops[i];
        int st = GTB_OK;
        switch (o.kind) {
        case OP_HD64:
            st = gtb_hori_diff_f64(&o.f[0], &o.f[1], &o.f[2], o.ni, o.nj, o.nk, o.stream);
            break;
        case OP_HD32:
            st = gtb_hori_diff_f32(&o.f[0], &o.f[1], &o.f[2], o.ni, o.nj, o.nk, o.stream);
            break;
        case OP_VA64:
            st = gtb_vert_adv_f64(&o.f[0], &o.f[1], &o.f[2], &o.f[3], &o.f[4], o.scalar, o.ni, o.nj, o.nk, o.stream);
            break;
        case OP_VA32:
            st = gtb_vert_adv_f32(&o.f[0], &o.f[1], &o.f[2], &o.f[3], &o.f[4], (float)o.scalar, o.ni, o.nj, o.nk,
                o.stream);
            break;
        case OP_TRACERS:
            st = gtb_prepare_tracers_f64(o.outs.data(), o.ins.data(), (int)o.outs.size(), &o.f[0], o.ni, o.nj, o.nk,
                o.stream);
            break;
        case OP_HALO:
            st = gtb_halo_exchange(o.halo, o.ptrs.data(), (int)o.ptrs.size(), o.stream);
            break;
        case OP_SGATE:
            st = gtb_stencil_gate(o.gate_flag, o.gate_value, o.gate_post);
            break;
        case OP_HGATE:
            st = gtb_halo_gate(o.halo, o.gate_flag, o.gate_value);
            break;
        case OP_RECORD:
            GTB_CUDA(cudaEventRecord(s->events[o.event], as_stream(o.stream)));
            break;
        case OP_WAIT:
            GTB_CUDA(cudaStreamWaitEvent(as_stream(o.stream), s->events[o.event], 0));
            break;
        case OP_STAMP:
            st = gtb_stamp(o.gate_post, o.stream);
            break;
        case OP_MARK:
            GTB_CUDA(cudaEventRecord(s->marks[o.event], as_stream(o.stream)));
            break;
        }
        if (st)
            return st;
    }
    return GTB_OK;
}
