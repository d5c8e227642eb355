/*
 * RappasB200 -- Panama FFM (java.lang.foreign, JDK 22+) binding of librappas_b200.so: the GPU placement
 * engine behind the C ABI of include/rappas_b200.h, plus the loop that replaces the call
 *     asp.processQueries(fp, placements, ...)          (src/main_v2/Main_PLACEMENT_v07.java:255-257)
 *
 * SOURCE ONLY: not compiled or run where it is shipped (no JDK there).  What IS tested is the C ABI it
 * binds, through the ctypes mirror rappas_b200/_abi.py (same prototypes).
 *
 * The host keeps R1 (FASTA), R2 (duplicates) and R11 (jplace) or hands them to the library too:
 * rp_reads_load_fasta + rp_jplace_write reproduce FASTAPointer / the MD5 duplicate map / the row assembly
 * (PlacementProcess.java:568-629, 974-1047) natively.
 */
package core.algos;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

public final class RappasB200 implements AutoCloseable {
    private static final Linker L = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup("librappas_b200.so", Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return L.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }

    // int rp_db_load_file(const char* path, const int32_t* devices, int32_t n, int32_t partitioned, rp_db** out)
    private static final MethodHandle DB_LOAD_FILE = fn("rp_db_load_file", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle DB_FREE = fn("rp_db_free", FunctionDescriptor.ofVoid(ADDRESS));
    // int rp_place_batch(rp_db*, const rp_place_cfg*, const uint8_t* seq, const uint64_t* seq_off, int64_t n,
    //                    int32_t* n_rows, uint16_t* node, float* score, double* lwr, int32_t* counts, int32_t* status)
    private static final MethodHandle PLACE = fn("rp_place_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle READS_LOAD = fn("rp_reads_load_fasta", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle READS_FREE = fn("rp_reads_free", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle READS_DESCRIBE = fn("rp_reads_describe", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle READS_UNIQUE = fn("rp_reads_unique", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    // int rp_jplace_write(path, reads, K, n_rows, node, score, lwr, status, edge_id, branch_len, n_nodes, newick, invocation, guppy, not_placed, &n)
    private static final MethodHandle JPLACE = fn("rp_jplace_write", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("rp_last_error", FunctionDescriptor.of(ADDRESS));
    // int rp_host_alloc(void** out, uint64_t bytes); void rp_host_free(void*): page-locked batch buffers, so that the
    // library's H2D / kernel / D2H pipeline copies straight from / to them (a pageable buffer is staged by the library)
    private static final MethodHandle HOST_ALLOC = fn("rp_host_alloc", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle HOST_FREE = fn("rp_host_free", FunctionDescriptor.ofVoid(ADDRESS));

    /** A reusable page-locked buffer of `bytes` bytes (free with freePinned). */
    public MemorySegment allocPinned(long bytes) throws Throwable {
        MemorySegment pp = arena.allocate(ADDRESS);
        check((int) HOST_ALLOC.invokeExact(pp, bytes));
        return pp.get(ADDRESS, 0).reinterpret(bytes);
    }

    public void freePinned(MemorySegment m) throws Throwable { HOST_FREE.invokeExact(m); }

    /** struct rp_place_cfg { int32 keep_at_most; float keep_factor; int32 treat_amb; int32 amb_with_max; float ns_bound; int32 reserved0; } */
    static final MemoryLayout CFG = MemoryLayout.structLayout(JAVA_INT, JAVA_FLOAT, JAVA_INT, JAVA_INT, JAVA_FLOAT, JAVA_INT);

    private final Arena arena = Arena.ofConfined();
    private final MemorySegment db;

    private static void check(int rc) throws Throwable {
        if (rc != 0) {
            MemorySegment msg = ((MemorySegment) LAST_ERROR.invokeExact()).reinterpret(512);
            throw new IllegalStateException("rappas_b200 error " + rc + ": " + msg.getString(0));
        }
    }

    /** Loads the .rgdb written by RgdbExporter onto the given CUDA devices (replicated). */
    public RappasB200(String rgdbPath, int[] devices) throws Throwable {
        MemorySegment out = arena.allocate(ADDRESS);
        check((int) DB_LOAD_FILE.invokeExact(arena.allocateFrom(rgdbPath), arena.allocateFrom(JAVA_INT, devices), devices.length, 0, out));
        db = out.get(ADDRESS, 0);
    }

    /**
     * The whole placement of one query file: FASTA -> unique sequences -> GPU -> .jplace, i.e. what
     * Main_PLACEMENT_v07.java:246-315 does around processQueries.  edgeId/branchLen are indexed by the node ids
     * of session.originalTree (PhyloNode.getJplaceEdgeId() / getBranchLengthToAncestor()).
     */
    public long placeFile(String fastaPath, String jplacePath, String notPlacedPath, int keepAtMost, float keepFactor,
                          boolean treatAmbiguities, boolean ambWithMax, float nsBound, boolean guppyCompatible,
                          int[] edgeId, float[] branchLen, String newick, String invocation) throws Throwable {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment pr = a.allocate(ADDRESS);
            check((int) READS_LOAD.invokeExact(a.allocateFrom(fastaPath), pr));
            MemorySegment reads = pr.get(ADDRESS, 0);
            try {
                MemorySegment nRec = a.allocate(JAVA_LONG), nUniq = a.allocate(JAVA_LONG), nGrp = a.allocate(JAVA_LONG);
                check((int) READS_DESCRIBE.invokeExact(reads, nRec, nUniq, nGrp));
                long n = nUniq.get(JAVA_LONG, 0);
                MemorySegment pSeq = a.allocate(ADDRESS), pOff = a.allocate(ADDRESS);
                check((int) READS_UNIQUE.invokeExact(reads, pSeq, pOff));
                MemorySegment cfg = a.allocate(CFG);
                cfg.set(JAVA_INT, 0, keepAtMost); cfg.set(JAVA_FLOAT, 4, keepFactor);
                cfg.set(JAVA_INT, 8, treatAmbiguities ? 1 : 0); cfg.set(JAVA_INT, 12, ambWithMax ? 1 : 0);
                cfg.set(JAVA_FLOAT, 16, nsBound); cfg.set(JAVA_INT, 20, 0);
                MemorySegment nRows = a.allocate(JAVA_INT, n), status = a.allocate(JAVA_INT, n), counts = a.allocate(JAVA_INT, 4 * n);
                MemorySegment node = a.allocate(JAVA_SHORT, n * keepAtMost), score = a.allocate(JAVA_FLOAT, n * keepAtMost),
                              lwr = a.allocate(JAVA_DOUBLE, n * keepAtMost);
                check((int) PLACE.invokeExact(db, cfg, pSeq.get(ADDRESS, 0), pOff.get(ADDRESS, 0), n, nRows, node, score, lwr, counts, status));
                // a status 3 (unsupported character) is where the reference prints "Unsupported state" and exits (AmbigSequenceKnife.java:124-128)
                MemorySegment nPl = a.allocate(JAVA_LONG);
                check((int) JPLACE.invokeExact(a.allocateFrom(jplacePath), reads, keepAtMost, nRows, node, score, lwr, status,
                        a.allocateFrom(JAVA_INT, edgeId), a.allocateFrom(JAVA_FLOAT, branchLen), edgeId.length,
                        a.allocateFrom(newick), a.allocateFrom(invocation), guppyCompatible ? 1 : 0,
                        notPlacedPath == null ? MemorySegment.NULL : a.allocateFrom(notPlacedPath), nPl));
                return nPl.get(JAVA_LONG, 0);
            } finally {
                READS_FREE.invokeExact(reads);
            }
        }
    }

    @Override public void close() {
        try { DB_FREE.invokeExact(db); } catch (Throwable t) { throw new RuntimeException(t); }
        arena.close();
    }
}
