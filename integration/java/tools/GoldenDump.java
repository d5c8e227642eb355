/*
 * GoldenDump -- runs the UNMODIFIED RAPPAS placement (PlacementProcess.processQueries,
 * src/core/algos/PlacementProcess.java:471-1118, with its own AmbigSequenceKnife, CustomHash_v4_FastUtil81 and
 * json-simple row assembly) on a .rgdb DB + a FASTA file and writes what it placed as one JSON file:
 *
 *   { "k":..,"n_nodes":..,"keep_at_most":..,"keep_factor":..,
 *     "edge_of_node":[..],                       // PhyloTree.getJplaceMappingNodeIdToJP per node id
 *     "placements":[ {"nm":[..], "p":[[edge_num, likelihood, like_weight_ratio, distal_length, pendant_length],..]}, ..],
 *     "elapsed_ms": ..  }                        // the "placement execution took" figure of Main_PLACEMENT_v07.java:325
 *
 * tests/test_golden_jvm.py (activated by files under tests/golden_jvm/) holds the CPU oracle and the CUDA library
 * to these rows: likelihood bit for bit, like_weight_ratio to 1e-9, edge identical except among exactly tied
 * scores.  That is the pin of SURVEY.md 8c the repository cannot produce itself (no JDK / fastutil in its images).
 *
 * The session is assembled from public members only (SessionNext_v2.java:43-110): states, k, thresholds, the hash
 * from RgdbImporter, and a caterpillar tree with exactly n_nodes nodes so that originalTree.getById(x) exists for
 * every node id of the DB (PlacementProcess.java:495-496, 1004-1047).  onlyFakes = true: no remap (:866-916).
 *
 * SOURCE ONLY -- never compiled here (see RgdbImporter).  usage:
 *   java -cp RAPPAS.jar:fastutil-8.2.2.jar:json_simple-1.1.jar:. tools.GoldenDump db.rgdb reads.fasta out.json [keepAtMost keepFactor ambWithMax]
 */
package tools;

import core.algos.AmbigSequenceKnife;
import core.algos.ISequenceKnife;
import core.algos.PlacementProcess;
import core.algos.SequenceKnife;
import inputs.FASTAPointer;
import main_v2.SessionNext_v2;
import org.json.simple.JSONArray;
import org.json.simple.JSONObject;
import tree.NewickReader;
import tree.PhyloTree;

import java.io.BufferedWriter;
import java.io.File;
import java.io.FileWriter;
import java.nio.file.Paths;

public final class GoldenDump {
    /** rooted caterpillar with n nodes (n odd: (n+1)/2 leaves): ((((t0,t1),t2),t3),...); every edge length 0.1 */
    static String caterpillar(int nNodes) {
        int leaves = (nNodes + 1) / 2;
        StringBuilder sb = new StringBuilder("(t0:0.1,t1:0.1)");
        for (int i = 2; i < leaves; i++) sb.insert(0, '(').append(":0.1,t").append(i).append(":0.1)");
        return sb.append(';').toString();
    }

    public static void main(String[] a) throws Exception {
        RgdbImporter db = RgdbImporter.load(Paths.get(a[0]));
        int keepAtMost = a.length > 3 ? Integer.parseInt(a[3]) : 7;
        float keepFactor = a.length > 4 ? Float.parseFloat(a[4]) : 0.01f;
        boolean ambWithMax = a.length > 5 && Boolean.parseBoolean(a[5]);
        SessionNext_v2 s = new SessionNext_v2(db.k, db.k, 1.5f, 1, Float.MIN_VALUE, db.thrLin, db.thrLog10);
        s.associateStates(db.states);
        s.associateHash(db.hash, true);
        s.originalTree = NewickReader.parseNewickTree2(caterpillar(db.nNodes), true, false);
        s.originalTree.initIndexes();
        s.originalTree.resetJplaceEdgeIds();
        if (s.originalTree.getNodeCount() != db.nNodes)
            throw new IllegalStateException("tree has " + s.originalTree.getNodeCount() + " nodes, DB says " + db.nNodes);
        FASTAPointer fp = new FASTAPointer(new File(a[1]), false);            // gaps kept: Main_PLACEMENT_v07.java:195
        JSONArray placements = new JSONArray();
        File tmp = File.createTempFile("golden", ".tsv");
        BufferedWriter tsv = new BufferedWriter(new FileWriter(tmp)), notPlaced = new BufferedWriter(new FileWriter(tmp + ".np"));
        PlacementProcess asp = new PlacementProcess(s, Float.NEGATIVE_INFINITY, Integer.MAX_VALUE);
        ISequenceKnife sk = new AmbigSequenceKnife(s.k, s.minK, s.states, SequenceKnife.SAMPLING_LINEAR);   // :112
        long t0 = System.currentTimeMillis();
        asp.processQueries(fp, placements, tsv, notPlaced, sk, 0, tmp.getParentFile(), keepAtMost, keepFactor, false, true, ambWithMax);
        long t1 = System.currentTimeMillis();
        tsv.close(); notPlaced.close(); fp.closePointer();
        JSONObject out = new JSONObject();
        out.put("k", db.k); out.put("n_nodes", db.nNodes); out.put("keep_at_most", keepAtMost); out.put("keep_factor", keepFactor);
        out.put("amb_with_max", ambWithMax); out.put("elapsed_ms", t1 - t0);
        JSONArray edges = new JSONArray();
        for (int x = 0; x < db.nNodes; x++) edges.add(s.originalTree.getJplaceMappingNodeIdToJP(x));
        out.put("edge_of_node", edges);
        out.put("placements", placements);
        try (FileWriter w = new FileWriter(a[2])) { w.write(out.toJSONString()); }
        System.out.println("placed " + placements.size() + " distinct sequences in " + (t1 - t0) + " ms");
    }
}
