/*
 * RgdbExporter -- writes the phylo-kmer hash of a loaded RAPPAS session as the flat .rgdb file that
 * librappas_b200 loads (rp_db_load_file; layout in DESIGN.md section 2).
 *
 * SOURCE ONLY: written against RAPPAS' classes (SessionNext_v2, CustomHash_v4_FastUtil81, fastutil 8.2.2),
 * NOT compiled or run in the repository that ships it -- that environment has no JDK.  It follows the
 * walk of SessionNext_v2.saveToJSON (src/main_v2/SessionNext_v2.java:250-261).
 *
 * Usage from Main_DBBUILD_3 after the session is stored, or from a small CLI:
 *     RgdbExporter.export(session, Paths.get(dbPath + ".rgdb"));
 */
package core.hash;

import core.DNAStatesShifted;
import it.unimi.dsi.fastutil.chars.Char2FloatMap;
import it.unimi.dsi.fastutil.chars.Char2FloatOpenHashMap;
import it.unimi.dsi.fastutil.objects.Object2ObjectMap;
import main_v2.SessionNext_v2;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.channels.FileChannel;
import java.nio.file.Path;
import java.nio.file.StandardOpenOption;

public final class RgdbExporter {
    private RgdbExporter() {}

    private static long align64(long x) { return (x + 63L) & ~63L; }

    /** k-mer code of the ABI: nucleotides = compressMer bytes read little-endian; amino acids = sum b_i * 32^i. */
    static long code(byte[] key, boolean nucl) {
        long c = 0;
        if (nucl) for (int b = 0; b < key.length; b++) c |= (key[b] & 0xFFL) << (8 * b);
        else      for (int b = 0; b < key.length; b++) c |= ((long) key[b]) << (5 * b);
        return c;
    }

    public static void export(SessionNext_v2 s, Path out) throws IOException {
        final var map = s.hash.getHash();  // Object2ObjectOpenCustomHashMap<byte[],Char2FloatOpenHashMap>
        final boolean nucl = s.states instanceof DNAStatesShifted;
        long nKeys = map.size(), nPost = 0;
        for (Object2ObjectMap.Entry<byte[], Char2FloatOpenHashMap> e : map.object2ObjectEntrySet()) nPost += e.getValue().size();
        final long offKeys = 128, offOffsets = align64(offKeys + 8 * nKeys),
                   offNodes = align64(offOffsets + 8 * (nKeys + 1)), offScores = align64(offNodes + 2 * nPost);
        try (FileChannel ch = FileChannel.open(out, StandardOpenOption.CREATE, StandardOpenOption.WRITE,
                                               StandardOpenOption.TRUNCATE_EXISTING)) {
            ByteBuffer h = ByteBuffer.allocate(80).order(ByteOrder.LITTLE_ENDIAN);
            h.put(new byte[]{'R', 'G', 'D', 'B', 0, 0, 0, 1});
            // alphabet: 0 nucleotides, 1 amino acids, 2 amino acids of a DB built with --convertUO (AAStates.java:118-123:
            // U and O are then states -- C and L -- and queries may contain them; RP_ALPHA_AMINO_UO)
            int alphabet = nucl ? 0 : 1;
            if (!nucl) {
                try { s.states.stateToByte('U'); alphabet = 2; } catch (Exception notConverted) { /* plain AAStates(false) */ }
            }
            h.putInt(alphabet).putInt(s.k).putInt(s.originalTree.getNodeCount());
            h.putFloat(s.PPStarThresholdAsLog10).putFloat(s.PPStarThreshold).putInt(0);
            h.putLong(nKeys).putLong(nPost).putLong(offKeys).putLong(offOffsets).putLong(offNodes).putLong(offScores);
            h.flip();
            ch.write(h, 0);
            // four positional streams, one pass over the map; postings in char2FloatEntrySet() order
            final int CH = 1 << 20;
            ByteBuffer keys = ByteBuffer.allocate(8 * CH).order(ByteOrder.LITTLE_ENDIAN);
            ByteBuffer offs = ByteBuffer.allocate(8 * CH).order(ByteOrder.LITTLE_ENDIAN);
            ByteBuffer nodes = ByteBuffer.allocate(2 * CH).order(ByteOrder.LITTLE_ENDIAN);
            ByteBuffer scores = ByteBuffer.allocate(4 * CH).order(ByteOrder.LITTLE_ENDIAN);
            long pKeys = offKeys, pOffs = offOffsets, pNodes = offNodes, pScores = offScores, posting = 0;
            for (Object2ObjectMap.Entry<byte[], Char2FloatOpenHashMap> e : map.object2ObjectEntrySet()) {
                if (!keys.hasRemaining()) { keys.flip(); pKeys += ch.write(keys, pKeys); keys.clear(); }
                if (!offs.hasRemaining()) { offs.flip(); pOffs += ch.write(offs, pOffs); offs.clear(); }
                keys.putLong(code(e.getKey(), nucl));
                offs.putLong(posting);
                for (Char2FloatMap.Entry pe : e.getValue().char2FloatEntrySet()) {
                    if (!nodes.hasRemaining()) { nodes.flip(); pNodes += ch.write(nodes, pNodes); nodes.clear(); }
                    if (!scores.hasRemaining()) { scores.flip(); pScores += ch.write(scores, pScores); scores.clear(); }
                    nodes.putShort((short) pe.getCharKey());
                    scores.putFloat(pe.getFloatValue());
                    posting++;
                }
            }
            if (!offs.hasRemaining()) { offs.flip(); pOffs += ch.write(offs, pOffs); offs.clear(); }
            offs.putLong(posting);
            keys.flip(); ch.write(keys, pKeys);
            offs.flip(); ch.write(offs, pOffs);
            nodes.flip(); ch.write(nodes, pNodes);
            scores.flip(); ch.write(scores, pScores);
        }
    }
}
