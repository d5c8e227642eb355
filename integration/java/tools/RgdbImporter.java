/*
 * RgdbImporter -- the inverse of RgdbExporter: reads a flat .rgdb file (layout: DESIGN.md section 2,
 * rappas_b200/csrc/rp_db.cu) into RAPPAS' own CustomHash_v4_FastUtil81 through its own addTuple
 * (src/core/hash/CustomHash_v4_FastUtil81.java:73-90), so that the UNMODIFIED reference can place reads against
 * the very DB the GPU library and the CPU oracle are tested on.  Used by GoldenDump.
 *
 * SOURCE ONLY: written against RAPPAS v1.22 + fastutil 8.2.2; neither a JDK nor the fastutil jar exists in the
 * repository's build image or on its GPU box (probed: `java` not found), so this has never been compiled.
 * tools/make_jvm_golden.sh compiles and runs it on a machine that has both.
 */
package tools;

import core.AAStates;
import core.DNAStatesShifted;
import core.States;
import core.hash.CustomHash_v4_FastUtil81;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.channels.FileChannel;
import java.nio.file.Path;
import java.nio.file.StandardOpenOption;

public final class RgdbImporter {
    public int alphabet, k, nNodes;
    public float thrLog10, thrLin;
    public long nKeys, nPostings;
    public States states;
    public CustomHash_v4_FastUtil81 hash;

    /** ABI k-mer code -> the byte[] RAPPAS keys its hash on: compressMer bytes (nucl) or the raw state bytes (amino). */
    static byte[] word(long code, int alphabet, int k) {
        if (alphabet == 0) {
            byte[] states = new byte[k];
            for (int i = 0; i < k; i++) states[i] = (byte) ((code >>> (2 * i)) & 3L);
            return new DNAStatesShifted().compressMer(states);   // DNAStatesShifted.java:115-143
        }
        byte[] w = new byte[k];
        for (int i = 0; i < k; i++) w[i] = (byte) ((code >>> (5 * i)) & 31L);
        return w;                                                 // AAStates.compressMer is the identity (:195-197)
    }

    public static RgdbImporter load(Path file) throws IOException {
        RgdbImporter r = new RgdbImporter();
        try (FileChannel ch = FileChannel.open(file, StandardOpenOption.READ)) {
            ByteBuffer h = ByteBuffer.allocate(80).order(ByteOrder.LITTLE_ENDIAN);
            ch.read(h, 0);
            h.flip();
            byte[] magic = new byte[8];
            h.get(magic);
            if (magic[0] != 'R' || magic[1] != 'G' || magic[2] != 'D' || magic[3] != 'B') throw new IOException("not an .rgdb file");
            r.alphabet = h.getInt(); r.k = h.getInt(); r.nNodes = h.getInt();
            r.thrLog10 = h.getFloat(); r.thrLin = h.getFloat(); h.getInt();
            r.nKeys = h.getLong(); r.nPostings = h.getLong();
            long offKeys = h.getLong(), offOffsets = h.getLong(), offNodes = h.getLong(), offScores = h.getLong();
            r.states = r.alphabet == 0 ? new DNAStatesShifted() : new AAStates(r.alphabet == 2);
            r.hash = new CustomHash_v4_FastUtil81(r.k, r.states, CustomHash_v4_FastUtil81.NODES_UNION);
            ByteBuffer keys = ch.map(FileChannel.MapMode.READ_ONLY, offKeys, 8 * r.nKeys).order(ByteOrder.LITTLE_ENDIAN);
            ByteBuffer offs = ch.map(FileChannel.MapMode.READ_ONLY, offOffsets, 8 * (r.nKeys + 1)).order(ByteOrder.LITTLE_ENDIAN);
            for (long i = 0; i < r.nKeys; i++) {
                byte[] w = word(keys.getLong((int) (8 * i)), r.alphabet, r.k);
                long p0 = offs.getLong((int) (8 * i)), p1 = offs.getLong((int) (8 * (i + 1)));
                ByteBuffer nodes = ByteBuffer.allocate((int) (2 * (p1 - p0))).order(ByteOrder.LITTLE_ENDIAN);
                ByteBuffer scores = ByteBuffer.allocate((int) (4 * (p1 - p0))).order(ByteOrder.LITTLE_ENDIAN);
                ch.read(nodes, offNodes + 2 * p0);
                ch.read(scores, offScores + 4 * p0);
                // file order = insertion order; fastutil's own slot order then decides the iteration (and tie) order
                for (int p = 0; p < (int) (p1 - p0); p++)
                    r.hash.addTuple(w, scores.getFloat(4 * p), nodes.getShort(2 * p) & 0xFFFF, 0);
            }
        }
        return r;
    }
}
