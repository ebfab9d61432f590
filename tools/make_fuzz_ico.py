"""Random (seeded) unstructured stencils for differential tests of the b200-ico emitter against the
naive-ico oracle: stages on random location types, reductions over random chains with weights, sparse
dimensions, nested reductions, neighbour loops writing sparse fields, conditionals, horizontal-only and
vertical-only fields, level offsets.  Writes stencils/fuzzico_<nn>.iir.

Validity rules kept: a stage never reads through a chain (or at a neighbour) what it writes itself;
level offsets stay inside [k_start + 1, k_end - 1] intervals and never touch a field the (Parallel)
multistage writes; chains are taken from the table the
reference's ICOChainSize knows (dawn/src/dawn/CodeGen/IcoChainSizes.cpp)."""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from iir_builder import FULL, IIRBuilder, Interval, call, lit, loop_over, reduce_over, where  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "stencils")
LOCS = ["Cell", "Edge", "Vertex"]
CHAINS = {
    "Cell": [["Cell", "Edge"], ["Cell", "Vertex"], ["Cell", "Edge", "Cell"], ["Cell", "Edge", "Cell", "Edge", "Cell"]],
    "Edge": [["Edge", "Cell"], ["Edge", "Vertex"], ["Edge", "Cell", "Vertex"], ["Edge", "Cell", "Edge"]],
    "Vertex": [["Vertex", "Cell"], ["Vertex", "Edge"], ["Vertex", "Edge", "Vertex"]],
}
SIZE = {("Cell", "Edge"): 3, ("Cell", "Vertex"): 3, ("Cell", "Edge", "Cell"): 3,
        ("Cell", "Edge", "Cell", "Edge", "Cell"): 9, ("Edge", "Cell"): 2, ("Edge", "Vertex"): 2,
        ("Edge", "Cell", "Vertex"): 4, ("Edge", "Cell", "Edge"): 4, ("Vertex", "Cell"): 6,
        ("Vertex", "Edge"): 6, ("Vertex", "Edge", "Vertex"): 6}


def make(seed):
    rng = random.Random(1000 + seed)
    b = IIRBuilder("fuzzico_%02d" % seed, grid="Unstructured")
    dense = {loc: [b.field("%s_in%d" % (loc[0].lower(), n), "k", loc) for n in range(2)] for loc in LOCS}
    hor = {loc: b.field("%s_hor" % loc[0].lower(), "", loc) for loc in LOCS if rng.random() < 0.3}
    vert = b.field("vert", "k") if rng.random() < 0.4 else None
    sparse = {}
    outs = []
    produced = []   # fields written by earlier stages of the multistage
    g = b.global_var("beta", "Double", 0.5) if rng.random() < 0.5 else None

    def sparse_of(chain):
        key = tuple(chain)
        if key not in sparse:
            sparse[key] = b.field("sp_" + "".join(c[0].lower() for c in chain), "k", sparse_chain=list(chain))
        return sparse[key]

    def center(loc, allow_k, written):
        """expression at the stage's own element"""
        cands = [f for f in dense[loc] if f not in written]
        f = rng.choice(cands)
        # level offsets only on fields nobody in this (Parallel) multistage writes
        x = f.at(rng.choice([0, 0, 1, -1]) if (allow_k and f not in produced) else 0)
        if loc in hor and rng.random() < 0.3:
            x = x + hor[loc].at()
        if vert is not None and rng.random() < 0.3:
            x = x * vert.at()
        if g is not None and rng.random() < 0.3:
            x = x * g
        return x

    def neighbour(chain, written, depth, allow_k):
        """rhs of a reduction over `chain`: values at the neighbour, sparse values, centre values"""
        tgt = chain[-1]
        cands = [f for f in dense[tgt] if f not in written]
        f = rng.choice(cands)
        # flagged or resolved by location type (chains that end where they start need the flag)
        if chain[0] == chain[-1] or rng.random() < 0.6:
            x = f.at(rng.choice([0, 0, 1]) if (allow_k and f not in produced) else 0, nbh=True)
        else:
            x = f.at(0)
        r = rng.random()
        if r < 0.3:
            x = x * sparse_of(chain).at()
        elif r < 0.5 and chain[0] != chain[-1]:
            x = x + center(chain[0], False, written)
        if depth < 2 and rng.random() < 0.25:
            inner = rng.choice(CHAINS[tgt])
            if len(inner) <= 3:
                x = x + reduce_over(inner, neighbour(inner, written, depth + 1, False))
        return x

    def reduction(loc, written, allow_k):
        chain = rng.choice(CHAINS[loc])
        n = SIZE[tuple(chain)]
        weights = None
        if rng.random() < 0.4:
            weights = [lit(rng.choice([1.0, -1.0, 0.5, 2.0])) for _ in range(n)]
            if rng.random() < 0.3:
                weights = [center(loc, False, written) * w for w in weights]
        op = rng.choice(["+", "+", "+", "*", "max", "min"]) if weights is None else "+"
        init = {"+": 0.0, "*": 1.0, "max": -1e30, "min": 1e30}[op]
        if rng.random() < 0.3:
            init = center(loc, False, written)
        return reduce_over(chain, neighbour(chain, written, 1, allow_k), init=init, op=op, weights=weights)

    stages = []
    n_st = rng.randint(1, 3)
    for s in range(n_st):
        loc = rng.choice(LOCS)
        allow_k = rng.random() < 0.3 and seed < 16
        iv = Interval("Start", "End", 1, -1) if allow_k else FULL
        o = b.field("out%d_%s" % (s, loc[0].lower()), "k", loc)
        outs.append(o)
        written = {o}
        stmts = []
        kind = rng.random()
        if kind < 0.55:
            e = reduction(loc, written, allow_k)
            if rng.random() < 0.5:
                e = e + center(loc, allow_k, written)
            stmts.append(b.assign(o, e))
        elif kind < 0.75:
            decl, v = b.local("acc", reduction(loc, written, allow_k))
            stmts += [decl, b.if_(v > center(loc, False, written),
                                  [b.assign(o, v * 2.0)], [b.assign(o, center(loc, allow_k, written))])]
        else:
            chain = rng.choice([c for c in CHAINS[loc] if len(c) <= 3])
            sp = b.field("spout%d" % s, "k", sparse_chain=chain)
            tgt = chain[-1]
            nb = rng.choice(dense[tgt]).at(0, nbh=True)
            body = nb * 2.0 - center(loc, False, written)
            if rng.random() < 0.4 and tgt != loc:
                inner = rng.choice([c for c in CHAINS[tgt] if len(c) == 2])
                body = body + reduce_over(inner, rng.choice(dense[inner[-1]]).at(0, nbh=True))
            stmts.append(loop_over(chain, b.assign(sp, body)))
            stmts.append(b.assign(o, center(loc, allow_k, written)))
        stages.append(b.stage(b.do(iv, *stmts), location=loc))
        dense[loc].append(o)       # later stages may consume it (forces kernel splits / forwarding)
        produced.append(o)
    if seed >= 16:
        # seeds 16..: later stages OVERWRITE input fields that earlier stages read - at the element or
        # gathered through a chain (write-after-read: the emitter must not fuse such a stage into the
        # launch that still gathers the old values), then consume the new values through a chain again
        inputs = {loc: dense[loc][:2] for loc in LOCS}
        for _ in range(rng.randint(1, 2)):
            loc = rng.choice(LOCS)
            f, other = rng.sample(inputs[loc], 2) if rng.random() < 0.5 else (inputs[loc][0], inputs[loc][1])
            stages.append(b.stage(b.do(FULL, b.assign(f, other.at() * 2.0 + f.at())), location=loc))
            chain = rng.choice([c for c in CHAINS[loc] if c[-1] == loc] or CHAINS[loc])
            tgt = chain[-1]
            o = b.field("post%d_%s" % (len(stages), loc[0].lower()), "k", loc)
            stages.append(b.stage(b.do(FULL, b.assign(o, reduce_over(chain, inputs[tgt][0].at(0, nbh=True)))),
                                  location=loc))
    b.stencil(b.multistage("Parallel", *stages))
    return b


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    os.makedirs(OUT, exist_ok=True)
    for seed in range(count):
        with open(os.path.join(OUT, "fuzzico_%02d.iir" % seed), "w") as f:
            f.write(make(seed).dumps())
    print("wrote %d unstructured fuzz stencils" % count)


if __name__ == "__main__":
    main()
