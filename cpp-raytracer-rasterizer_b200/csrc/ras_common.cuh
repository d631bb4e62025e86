// ras_common.cuh -- shared by the two rasteriser pipelines (ras_sortlast.cu: the default; ras_tiles.cu: screen tiles).
#pragma once

namespace b2r {

// Device-side counters of one pipeline (each keeps its own copy at the head of its scratch buffer).  The first eight
// words are per frame and are left clear by the last kernel of a frame; `sticky` is only cleared by the host after it
// has been reported (asynchronous draws, see ras_take_error).
struct RasCounters {
    unsigned nBig, bigRows, bigSamples, err, totalRefs, blocksDone, totalJobs, pad0;
    unsigned sticky;
    unsigned pad1[7];
};

}  // namespace b2r
