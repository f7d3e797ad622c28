## R glue for the B200 engine: the reference's exported functions, same signatures, bodies replaced
## by .Call()s into src/rglue.c -> libderfinder_b200.so.  Argument checks, warnings and messages are
## the reference's (cited by file:line of lcolladotor/derfinder v1.35.0) and run BEFORE any native
## call.  BiocParallel arguments (mc.cores, BPPARAM.custom) are accepted and ignored: the library
## drives the GPU itself and must not be forked into Snow workers (one CUDA context per process).
##
## Not executable in the build image of this repository (no R there); see INTEGRATION.md.

.dfb <- new.env(parent = emptyenv())
.engine <- function(device = getOption("derfinderB200.device", 0L)) {
    if (is.null(.dfb$ctx)) .dfb$ctx <- .Call(dfbR_create, as.integer(device))
    .dfb$ctx
}

## Rle column lists of a DataFrame
.rle_lists <- function(df) list(values = lapply(df, S4Vectors::runValue), lengths = lapply(df, S4Vectors::runLength))
.lgl_runs <- function(x) list(lengths = S4Vectors::runLength(x), first = isTRUE(S4Vectors::runValue(x)[1]))
.as_lgl_rle <- function(p) S4Vectors::Rle(rep_len(c(p[[2]], !p[[2]]), length(p[[1]])), p[[1]])
.cov_dataframe <- function(h, names, which = 0L) {
    info <- .Call(dfbR_cov_info, h)
    cols <- lapply(seq_len(info[2]) - 1L, function(s) S4Vectors::Rle(.Call(dfbR_cov_vector, h, which, s)))
    names(cols) <- names
    S4Vectors::DataFrame(cols, check.names = FALSE)
}

## ---- filterData (R/filterData.R:89-241)
filterData <- function(data, cutoff = NULL, index = NULL, filter = "one", totalMapped = NULL, targetSize = 80e6, ...) {
    stopifnot(filter %in% c("one", "mean"))                                   # :104
    returnMean <- derfinder:::.advanced_argument("returnMean", FALSE, ...)
    returnCoverage <- derfinder:::.advanced_argument("returnCoverage", TRUE, ...)
    if (is.list(data) && !is(data, "DataFrame")) data <- S4Vectors::DataFrame(data, check.names = FALSE)
    ## library-size normalisation only when asked for AND targetSize != 0 (:119); totalMapped is recycled over the
    ## samples exactly like the reference's mapply(function(x, d) x / d, data, mappedPerXM) (:120-126)
    mapped <- NULL
    if (!is.null(totalMapped) && targetSize != 0) mapped <- rep_len(as.numeric(totalMapped) / targetSize, ncol(data))
    r <- .rle_lists(data)
    prev <- if (!is.null(index)) .lgl_runs(index)
    kind <- if (is.null(cutoff)) 0L else if (filter == "one") 1L else 2L
    h <- .Call(dfbR_filter, .engine(), r$values, r$lengths, nrow(data), mapped, kind,
               if (is.null(cutoff)) 0 else as.numeric(cutoff), prev$lengths, prev$first)
    res <- list(coverage = if (returnCoverage) .cov_dataframe(h, names(data)) else NULL,
                ## cutoff = NULL: nothing is filtered and `position` is `index`, NULL included (:133-135)
                position = if (is.null(cutoff)) index else .as_lgl_rle(.Call(dfbR_cov_position, h)))
    if (returnMean) res$meanCoverage <- S4Vectors::Rle(.Call(dfbR_cov_vector, h, 2L, 0L))
    attr(res, "dfb_cov") <- h      # device-resident copy: later stages skip the re-upload
    res
}

## ---- preprocessCoverage (R/preprocessCoverage.R:97-260)
preprocessCoverage <- function(coverageInfo, groupInfo = NULL, cutoff = 5, colsubset = NULL, lowMemDir = NULL, ...) {
    scalefac <- derfinder:::.advanced_argument("scalefac", 32, ...)
    chunksize <- derfinder:::.advanced_argument("chunksize", 5e6, ...)
    mc.cores <- derfinder:::.advanced_argument("mc.cores", getOption("mc.cores", 1L), ...)
    stopifnot(is.list(coverageInfo), c("coverage", "position") %in% names(coverageInfo))   # :131-134
    stopifnot(scalefac >= 0)
    if (is.null(coverageInfo$position)) stop("'position' must not be NULL")                # :163-165
    h <- attr(coverageInfo, "dfb_cov")
    means <- coverageInfo$meanCoverage
    if (!is.null(colsubset)) {                                                             # :145-160
        sub <- filterData(coverageInfo$coverage[, colsubset, drop = FALSE], cutoff = cutoff,
                          index = coverageInfo$position, returnMean = TRUE, ...)
        h <- attr(sub, "dfb_cov"); coverageInfo <- sub
        means <- sub$meanCoverage            # the re-filtered mean replaces a pre-existing one (:157-159)
    } else if (is.null(h)) {
        r <- .rle_lists(coverageInfo$coverage); p <- .lgl_runs(coverageInfo$position)
        h <- .Call(dfbR_cov_from_rle, .engine(), r$values, r$lengths, nrow(coverageInfo$coverage), p$lengths, p$first)
    }
    groups <- if (!is.null(groupInfo)) { stopifnot(is.factor(groupInfo)); as.integer(groupInfo) - 1L }
    .Call(dfbR_preprocess, .engine(), h, scalefac, groups, if (is.null(groupInfo)) 0L else nlevels(groupInfo))
    N <- .Call(dfbR_cov_info, h)[1]
    lens <- .Call(dfbR_chunk_lengths, N, if (is.null(chunksize)) 0 else chunksize, as.integer(mc.cores))
    k <- length(lens)
    split.idx <- S4Vectors::Rle(seq_len(k), lens)
    res <- list(
        coverageProcessed = if (is.null(lowMemDir)) .cov_dataframe(h, names(coverageInfo$coverage), 1L) else NULL,
        mclapplyIndex = if (is.null(lowMemDir)) lapply(seq_len(k), function(x) split.idx == x) else seq_len(k),
        position = coverageInfo$position,
        meanCoverage = if (is.null(means)) S4Vectors::Rle(.Call(dfbR_cov_vector, h, 2L, 0L)) else means,  # :191-193
        groupMeans = if (!is.null(groupInfo)) structure(lapply(seq_len(nlevels(groupInfo)) - 1L, function(g)
            S4Vectors::Rle(.Call(dfbR_cov_vector, h, 3L, g))), names = levels(groupInfo)) else list())
    attr(res, "dfb_cov") <- h
    attr(res, "dfb_means") <- list(meanCoverage = res$meanCoverage, groupMeans = res$groupMeans)
    res
}

## ---- calculateStats (R/calculateStats.R:70-132)
calculateStats <- function(coveragePrep, models, lowMemDir = NULL, ...) {
    stopifnot(length(intersect(names(coveragePrep), c("coverageProcessed", "mclapplyIndex", "position"))) == 3)  # :93-96
    stopifnot(length(intersect(names(models), c("mod", "mod0"))) == 2)                                           # :97-99
    adjustF <- derfinder:::.advanced_argument("adjustF", 0, ...)
    h <- attr(coveragePrep, "dfb_cov")
    if (is.null(h)) stop("coveragePrep does not carry a device handle: run preprocessCoverage() from this package")
    f <- .Call(dfbR_fstats, .engine(), h, models$mod, models$mod0, adjustF)
    res <- S4Vectors::Rle(f)
    attr(res, "dfb_cov") <- h   # marks these statistics as the ones cached on the device
    res
}

.cutoff_pair <- function(cutoff) {                     # R/findRegions.R:377-381
    if (length(cutoff) == 1) c(-cutoff, cutoff) else sort(cutoff)
}

.regions_granges <- function(x, chr) {
    nUp <- x$nUp; R <- length(x$value)
    gr <- GenomicRanges::GRanges(seqnames = S4Vectors::Rle(chr, R),
        ranges = IRanges::IRanges(start = x$start, end = x$end), strand = S4Vectors::Rle("*", R),
        value = x$value, area = x$area, indexStart = x$indexStart, indexEnd = x$indexEnd,
        cluster = S4Vectors::Rle(x$cluster), clusterL = S4Vectors::Rle(x$clusterL))
    names(gr) <- c(rep("up", nUp), rep("dn", R - nUp))                                   # R/findRegions.R:284
    gr
}

## ---- findRegions (R/findRegions.R:116-296), smooth = FALSE
findRegions <- function(position = NULL, fstats, chr, oneTable = TRUE, maxClusterGap = 300L,
    cutoff = quantile(fstats, 0.99, na.rm = TRUE), segmentIR = NULL, smooth = FALSE, weights = NULL,
    smoothFunction = bumphunter::locfitByCluster, ...) {
    basic <- derfinder:::.advanced_argument("basic", FALSE, ...)
    maxRegionGap <- derfinder:::.advanced_argument("maxRegionGap", 0L, ...)
    if (maxClusterGap < maxRegionGap)
        warning("'maxClusterGap' is less than 'maxRegionGap' which nullifies it's intended use.")   # :137-139
    if (!basic) {
        if (is.null(segmentIR) | smooth) stopifnot(!is.null(position))                   # :141-144
        if (smooth) stop("smooth = TRUE is not supported by the B200 engine")
    } else if (smooth) warning("Ignoring 'smooth' = TRUE since 'basic' = TRUE")          # :152
    if (is.null(segmentIR)) {
        if (!any(position)) { warning("Found no regions"); return(NULL) }                # :163-166
        p <- .lgl_runs(position); gap <- as.integer(maxRegionGap)
    } else if (!is.null(position)) {
        ## segmentIR is the cache of .clusterMakerRle(position, maxRegionGap, ranges = TRUE) (R/calculatePvalues.R:
        ## 234-237): the engine derives the same segments from `position`; check that the two belong together
        stopifnot(sum(IRanges::width(segmentIR)) == sum(position))
        p <- .lgl_runs(position); gap <- as.integer(maxRegionGap)
    } else {
        ## basic = TRUE with the segments only (index space, :155-170): a position index whose TRUE runs ARE the
        ## segments, one filtered-out base between neighbours
        w <- IRanges::width(segmentIR); k <- length(w)
        p <- list(lengths = as.integer(head(as.vector(rbind(w, 1L)), 2L * k - 1L)), first = TRUE); gap <- 0L
    }
    cp <- .cutoff_pair(as.numeric(cutoff))
    x <- .Call(dfbR_find_regions, .engine(), as.numeric(fstats), p$lengths, p$first, cp[1], cp[2],
               gap, as.integer(maxClusterGap), basic)
    if (is.null(x)) return(NULL)                                                         # no segments (:185-193)
    if (basic) return(S4Vectors::DataFrame(area = S4Vectors::Rle(x$area),
        width = S4Vectors::Rle(x$indexEnd - x$indexStart + 1L), stat = S4Vectors::Rle(x$value)))   # :274-278
    gr <- .regions_granges(x, chr)
    if (oneTable) gr else split(gr, names(gr))
}

## ---- calculatePvalues (R/calculatePvalues.R:157-430)
calculatePvalues <- function(coveragePrep, models, fstats, nPermute = 1L,
    seeds = as.integer(gsub("-", "", Sys.Date())) + seq_len(nPermute), chr,
    cutoff = quantile(fstats, 0.99, na.rm = TRUE), significantCut = c(0.05, 0.10), lowMemDir = NULL,
    smooth = FALSE, weights = NULL, smoothFunction = bumphunter::locfitByCluster, ...) {
    if (smooth) stop("smooth = TRUE is not supported by the B200 engine")
    stopifnot(length(intersect(names(coveragePrep), c("coverageProcessed", "mclapplyIndex", "position",
        "meanCoverage", "groupMeans"))) == 5)                                            # :181-184
    stopifnot(length(intersect(names(models), c("mod", "mod0"))) == 2)
    stopifnot(nPermute == length(seeds) | is.null(seeds))                                # :189
    stopifnot(length(significantCut) == 2 & all(significantCut >= 0 & significantCut <= 1))
    adjustF <- derfinder:::.advanced_argument("adjustF", 0, ...)
    maxRegionGap <- derfinder:::.advanced_argument("maxRegionGap", 0L, ...)
    maxClusterGap <- derfinder:::.advanced_argument("maxClusterGap", 300L, ...)
    h <- attr(coveragePrep, "dfb_cov")
    n <- nrow(models$mod)
    ## RNG stays in R: identical draws to the reference's loop (:315-318); NA / NULL seeds keep the stream
    perm <- vapply(seq_len(nPermute), function(i) {
        if (!is.null(seeds) && !is.na(seeds[i])) set.seed(seeds[i])
        sample(n) - 1L
    }, integer(n))
    dim(perm) <- c(n, nPermute)
    cp <- .cutoff_pair(as.numeric(cutoff))
    cached <- identical(attr(fstats, "dfb_cov"), h)       # statistics still on the device: no re-upload
    ## the reference summarises whatever meanCoverage / groupMeans vectors it is handed (:222-224, 253-262): the
    ## device copies are used only for the very objects preprocessCoverage() returned for this handle
    own <- attr(coveragePrep, "dfb_means")
    meanv <- if (!is.null(own) && identical(own$meanCoverage, coveragePrep$meanCoverage)) NULL
             else as.numeric(coveragePrep$meanCoverage)
    groupv <- if (!is.null(own) && identical(own$groupMeans, coveragePrep$groupMeans)) NULL
              else lapply(coveragePrep$groupMeans, as.numeric)
    x <- .Call(dfbR_calculate_pvalues, .engine(), h, if (cached) NULL else as.numeric(fstats), models$mod,
               models$mod0, adjustF, perm, cp[1], cp[2], as.integer(maxRegionGap), as.integer(maxClusterGap),
               length(coveragePrep$groupMeans), meanv, groupv)
    if (is.null(x)) return(list(regions = NULL, nullStats = NULL, nullWidths = NULL, nullPermutation = NULL))  # :245-251
    regs <- .regions_granges(x$regions, chr)
    regs$meanCoverage <- x$meanCoverage                                                  # :254-255
    gn <- names(coveragePrep$groupMeans)
    for (g in seq_along(gn)) S4Vectors::mcols(regs)[[paste0("mean", gn[g])]] <- x$groupMeans[[g]]          # :258-262
    for (g in seq_along(gn)[-1])                                                                           # :265-283
        S4Vectors::mcols(regs)[[paste0("log2FoldChange", gn[g], "vs", gn[1])]] <- log2(x$groupMeans[[g]] / x$groupMeans[[1]])
    regs$pvalues <- x$pvalues
    regs$significant <- factor(x$pvalues < significantCut[1], levels = c(TRUE, FALSE))   # :381-383
    if (length(x$nullStats) > 0) {
        qv <- qvalue::qvalue(x$pvalues)$qvalues                                          # stays in R (:390-399)
        regs$qvalues <- qv
        regs$significantQval <- factor(qv < significantCut[2], levels = c(TRUE, FALSE))
    } else {
        regs$qvalues <- NA; regs$significantQval <- NA                                   # :408-419
    }
    res <- list(regions = regs, nullStats = S4Vectors::Rle(x$nullStats), nullWidths = S4Vectors::Rle(x$nullWidths),
                nullPermutation = S4Vectors::Rle(x$nullPermutation))                     # :421-424
    ## device handles of this chromosome's regions and nulls: .mergeResultsGPU() pools them genome-wide
    attr(res, "dfb_regions") <- x$handles$regions
    attr(res, "dfb_nulls") <- x$handles$nulls
    res
}

## ---- the genome-wide pass of mergeResults (R/mergeResults.R:216-235, 300-307, 380-386) for chromosomes analysed in
## THIS R process (one or more GPUs-worth; with several processes, one per GPU, call .commInit() in each first):
## pooled p-values and FWER per chromosome, regions in the order calculatePvalues() returned them
.mergeResultsGPU <- function(chrResults, nPermute, cutoffFstatUsed, mergeArea = TRUE) {
    rh <- lapply(chrResults, attr, "dfb_regions"); nh <- lapply(chrResults, attr, "dfb_nulls")
    x <- .Call(dfbR_merge_results, .engine(), rh, nh, as.integer(nPermute), as.numeric(cutoffFstatUsed), isTRUE(mergeArea))
    names(x$pvalues) <- names(x$fwer) <- names(chrResults)
    x
}
## one process per GPU (e.g. one Rscript per device): rank 0 writes the 128-byte id to `idFile`, the others read it
.commInit <- function(world, rank, idFile) {
    if (rank == 0L) { id <- .Call(dfbR_comm_unique_id); writeBin(id, idFile) }
    else { while (!file.exists(idFile) || file.size(idFile) < 128) Sys.sleep(0.05); id <- readBin(idFile, "raw", 128L) }
    .Call(dfbR_comm_init_rank, .engine(), as.integer(world), as.integer(rank), id)
    invisible(TRUE)
}

## ---- regionMatrix per chromosome (R/regionMatrix.R:170-283), coverageMatrix only
.regionMatrixByChr <- function(covInfo, chr, cutoff, L, totalMapped, targetSize, runFilter = TRUE, maxRegionGap = 0L,
                               maxClusterGap = 300L) {
    if (runFilter) {                                                                     # :193-212
        filt <- filterData(covInfo$coverage, cutoff = cutoff, index = covInfo$position, filter = "mean",
                           totalMapped = totalMapped, targetSize = targetSize, returnMean = TRUE, returnCoverage = FALSE)
    } else {                                                                             # :213-228: already filtered
        filt <- covInfo
        if (is.null(attr(filt, "dfb_cov"))) {
            r <- .rle_lists(filt$coverage); p <- .lgl_runs(filt$position)
            attr(filt, "dfb_cov") <- .Call(dfbR_cov_from_rle, .engine(), r$values, r$lengths, nrow(filt$coverage),
                                           p$lengths, p$first)
        }
    }
    h <- attr(filt, "dfb_cov")
    regs <- findRegions(position = filt$position, fstats = filt$meanCoverage, chr = chr, cutoff = cutoff,
                        maxRegionGap = maxRegionGap, maxClusterGap = maxClusterGap)
    if (is.null(regs)) return(list(regions = GenomicRanges::GRanges(), coverageMatrix = NULL))             # :230-239
    if (!(length(L) %in% c(1L, length(covInfo$coverage)))) {                                               # :254-258
        warning("Invalid 'L' value so it won't be used. It has to either be a integer/numeric vector of length 1 or length equal to the number of samples.")
        L <- NULL
    }
    m <- .Call(dfbR_region_matrix, .engine(), h, regs$indexStart, regs$indexEnd, if (is.null(L)) NULL else as.numeric(L))
    colnames(m) <- names(covInfo$coverage)
    list(regions = regs, coverageMatrix = m)
}

## ---- regionMatrix (R/regionMatrix.R:124-168): chromosomes one after the other on the GPU (the reference's bpmapply
## over chromosomes :164 would fork the CUDA context); returnBP = TRUE is served by the reference's own
## getRegionCoverage() on the returned regions
regionMatrix <- function(fullCov, cutoff = 5, L, totalMapped = 80e6, targetSize = 80e6, runFilter = TRUE,
    returnBP = TRUE, ...) {
    stopifnot(!is.null(cutoff), !is.null(names(fullCov)))                                 # :128-131
    if (!runFilter) {                                                                     # :133-146
        if (any(totalMapped != targetSize)) stop("When using 'runFilter' = FALSE the arguments 'totalMapped' and 'targetSize' are not used. If you have not normalized the data, please do so before running regionMatrix(runFilter = FALSE).")
        ok <- vapply(fullCov, function(x) is.list(x) && all(c("coverage", "position", "meanCoverage") %in% names(x)), logical(1))
        if (!all(ok)) stop("When 'runFilter' = FALSE, all the elements of 'fullCov' are expected to be the output from filterData(..., returnMean=TRUE) with some non-NULL 'cutoff'.")
    }
    maxRegionGap <- derfinder:::.advanced_argument("maxRegionGap", 0L, ...)
    maxClusterGap <- derfinder:::.advanced_argument("maxClusterGap", 300L, ...)
    res <- lapply(names(fullCov), function(chr) {
        covInfo <- fullCov[[chr]]
        if (!is.list(covInfo) || is(covInfo, "DataFrame")) covInfo <- list(coverage = covInfo, position = NULL)
        out <- .regionMatrixByChr(covInfo, chr, cutoff, L, totalMapped, targetSize, runFilter, maxRegionGap, maxClusterGap)
        if (returnBP && length(out$regions) > 0)
            out$bpCoverage <- derfinder::getRegionCoverage(fullCov = fullCov[chr], regions = out$regions,
                totalMapped = totalMapped, targetSize = targetSize, verbose = FALSE)
        out
    })
    names(res) <- names(fullCov)
    res
}

## ---- railMatrix's per-chunk coverage matrix (.railMatChrRegion, R/railMatrix.R:238-262): the caller imports the
## selected ranges of every sample BigWig (`import(sampleFile, selection = reduce(regs), as = "RleList")`) and hands
## the Rle columns over; one call replaces the bpmapply over .railMatChrRegionCov (:264-290)
.railMatChrRegion <- function(regCovs, regs, mappedPerXM, L) {
    filt <- filterData(regCovs, cutoff = NULL, totalMapped = mappedPerXM, targetSize = 1, returnCoverage = FALSE)
    m <- .Call(dfbR_view_sums, .engine(), attr(filt, "dfb_cov"), GenomicRanges::start(regs), GenomicRanges::end(regs),
               as.numeric(L))
    colnames(m) <- names(regCovs)
    list(coverageMatrix = m)
}

## the same from RAW BigWig data sections (a reader that hands out the inflated leaf blocks instead of decoded
## RleLists, e.g. a thin wrapper around the file's R-tree): the decode runs on the device (dfb_cov_from_bigwig).
## sections: list of raw vectors, secOff: list of byte offsets (+ total), chromId: the files' ids of the chromosome;
## regs must be sorted and disjoint (GenomicRanges::reduce()d chunks, as railMatrix() passes them)
.railMatChrRegionBigWig <- function(sections, secOff, chromId, regs, mappedPerXM, L) {
    st <- GenomicRanges::start(regs); en <- GenomicRanges::end(regs)
    h <- .Call(dfbR_cov_from_bigwig, .engine(), sections, lapply(secOff, as.numeric), as.integer(chromId), st, en,
               if (is.null(mappedPerXM)) NULL else as.numeric(mappedPerXM))
    w <- en - st + 1L; off <- cumsum(c(0L, head(w, -1L)))                 # rows of region k: off[k] + 1 .. off[k] + w[k]
    m <- .Call(dfbR_view_sums, .engine(), h, off + 1L, off + w, as.numeric(L))
    colnames(m) <- names(sections)
    list(coverageMatrix = m)
}

.calcPval <- function(areas, nullareas) .Call(dfbR_calc_pval, .engine(), as.numeric(areas), as.numeric(nullareas))
.calculateFWER <- function(cutoffFstatUsed, areaNull, permutation, nPermute, areaReg)
    .Call(dfbR_calc_fwer, .engine(), cutoffFstatUsed, as.numeric(areaNull), as.integer(permutation),
          as.integer(nPermute), as.numeric(areaReg))


## ---- railMatrix (R/railMatrix.R:115-175): ERs of the mean BigWig per chromosome, then the region x sample matrix.
## BigWig I/O stays with rtracklayer (as in the reference); filter, segmentation, clusters and the per-sample
## sum(Views(cov / mappedPerXM, regions)) / L run on the device.  Chromosomes and sample files are walked serially
## (BPPARAM arguments are accepted and ignored: the CUDA context must not be forked).
railMatrix <- function(chrs, summaryFiles, sampleFiles, L, cutoff = NULL, maxClusterGap = 300L, totalMapped = NULL,
    targetSize = 40e6, file.cores = 1L, chrlens = NULL, ...) {
    stopifnot(length(chrs) == length(summaryFiles))                                      # :119
    stopifnot(!is.null(cutoff))                                                          # :125
    stopifnot(length(L) != 1 | length(L) != length(sampleFiles))                         # :126 (sic)
    if (!is.null(totalMapped)) stopifnot(length(totalMapped) != 1 | length(totalMapped) != length(sampleFiles))
    if (is.null(chrlens)) chrlens <- rep(list(NULL), length(chrs))
    mappedPerXM <- if (!is.null(totalMapped) & targetSize != 0) totalMapped / targetSize else NULL        # :151-155
    chunksize <- derfinder:::.advanced_argument("chunksize", 1000, ...)
    res <- mapply(function(chr, summaryFile, chrlen) {
        meanCov <- derfinder::loadCoverage(files = summaryFile, chr = chr, chrlen = chrlen)                # :181
        filteredMean <- filterData(meanCov$coverage, cutoff = cutoff, filter = "mean", returnMean = TRUE, ...)
        regs <- findRegions(position = filteredMean$position, fstats = filteredMean$meanCoverage, chr = chr,
                            maxClusterGap = maxClusterGap, cutoff = cutoff, ...)
        if (is.null(regs)) return(list(regions = GenomicRanges::GRanges(), coverageMatrix = NULL))        # :190-192
        names(regs) <- seq_len(length(regs))
        GenomeInfoDb::seqlengths(regs) <- length(meanCov$coverage[[1]])
        ## regions in chunks of `chunksize` (:203-216), every sample's BigWig imported for the chunk's ranges only
        chunks <- split(seq_len(length(regs)), ceiling(seq_len(length(regs)) / chunksize))
        mats <- lapply(chunks, function(ii) {
            rg <- regs[ii]
            covs <- lapply(sampleFiles, function(f)
                rtracklayer::import(f, selection = GenomicRanges::reduce(rg), as = "RleList")[[chr]])      # :271-284
            names(covs) <- if (is.null(names(sampleFiles))) sampleFiles else names(sampleFiles)
            .railMatChrRegion(S4Vectors::DataFrame(covs, check.names = FALSE), rg,
                              if (is.null(mappedPerXM)) NULL else rep_len(mappedPerXM, length(sampleFiles)),
                              rep_len(L, length(sampleFiles)))$coverageMatrix
        })
        list(regions = regs, coverageMatrix = do.call(rbind, mats))
    }, chrs, summaryFiles, chrlens, SIMPLIFY = FALSE)
    names(res) <- chrs
    res
}

## ---- analyzeChr (R/analyzeChr.R:108-310): the per-chromosome driver; every numeric stage is this package's, the
## annotation and the on-disk layout are the reference's
analyzeChr <- function(chr, coverageInfo, models, cutoffPre = 5, cutoffFstat = 1e-08, cutoffType = "theoretical",
    nPermute = 1, seeds = as.integer(gsub("-", "", Sys.Date())) + seq_len(nPermute), groupInfo, txdb = NULL,
    writeOutput = TRUE, runAnnotation = TRUE, lowMemDir = file.path(chr, "chunksDir"), smooth = FALSE, weights = NULL,
    smoothFunction = bumphunter::locfitByCluster, ...) {
    stopifnot(length(intersect(cutoffType, c("empirical", "theoretical", "manual"))) == 1)                 # :117-120
    stopifnot(is.factor(groupInfo))
    stopifnot(length(intersect(names(models), c("mod", "mod0"))) == 2)
    returnOutput <- derfinder:::.advanced_argument("returnOutput", !writeOutput, ...)
    timeinfo <- c(init = Sys.time())
    if (writeOutput) {
        dir.create(chr, showWarnings = FALSE, recursive = TRUE)
        optionsStats <- list(models = models, cutoffPre = cutoffPre, cutoffFstat = cutoffFstat, cutoffType = cutoffType,
            nPermute = nPermute, seeds = seeds, groupInfo = groupInfo, lowMemDir = lowMemDir, analyzeCall = match.call(),
            cutoffFstatUsed = NULL, smooth = smooth, smoothFunction = smoothFunction, weights = weights, ...)
    }
    ## the device keeps the chunks: lowMemDir is accepted for signature compatibility and not used (:167)
    prep <- preprocessCoverage(coverageInfo = coverageInfo, groupInfo = groupInfo, cutoff = cutoffPre, lowMemDir = NULL, ...)
    timeinfo <- c(timeinfo, setup = Sys.time())
    fstats <- calculateStats(coveragePrep = prep, models = models, ...)                                   # :187
    timeinfo <- c(timeinfo, calculateStats = Sys.time())
    ## F cutoff (.calcFstatCutoff, :312-328)
    cutoffFstatUsed <- switch(cutoffType,
        empirical = quantile(as.numeric(fstats), cutoffFstat, na.rm = TRUE),
        theoretical = { n <- nrow(models$mod); df1 <- ncol(models$mod); df0 <- ncol(models$mod0)
                        qf(cutoffFstat, df1 - df0, n - df1, lower.tail = FALSE) },
        manual = cutoffFstat)
    if (writeOutput) {
        optionsStats$cutoffFstatUsed <- cutoffFstatUsed
        save(optionsStats, file = file.path(chr, "optionsStats.Rdata"))
        save(fstats, file = file.path(chr, "fstats.Rdata"))
        coveragePrep <- prep; save(coveragePrep, file = file.path(chr, "coveragePrep.Rdata"))
    }
    regions <- calculatePvalues(coveragePrep = prep, models = models, fstats = fstats, nPermute = nPermute,
        seeds = seeds, chr = chr, cutoff = cutoffFstatUsed, smooth = smooth, weights = weights,
        smoothFunction = smoothFunction, ...)                                                              # :237
    timeinfo <- c(timeinfo, calculatePvalues = Sys.time())
    if (writeOutput) save(regions, file = file.path(chr, "regions.Rdata"))
    annotation <- NULL
    if (runAnnotation && !is.null(regions$regions)) {                                                      # :262-283
        annotation <- if (is.null(txdb)) bumphunter::matchGenes(x = regions$regions, subject = bumphunter::annotateTranscripts(
            TxDb.Hsapiens.UCSC.hg19.knownGene::TxDb.Hsapiens.UCSC.hg19.knownGene))
            else bumphunter::matchGenes(x = regions$regions, subject = bumphunter::annotateTranscripts(txdb))
        if (writeOutput) save(annotation, file = file.path(chr, "annotation.Rdata"))
    }
    timeinfo <- c(timeinfo, annotation = Sys.time())
    if (writeOutput) save(timeinfo, file = file.path(chr, "timeinfo.Rdata"))
    if (returnOutput) list(timeinfo = timeinfo, optionsStats = if (writeOutput) optionsStats else NULL,
        coveragePrep = prep, fstats = fstats, regions = regions, annotation = annotation) else invisible(NULL)
}

## ---- mergeResults (R/mergeResults.R:109-357): the genome-wide pass.  Same signature, same files
## (fullRegions.Rdata, fullNullSummary.Rdata, fullFstats.Rdata, fullTime.Rdata, optionsMerge.Rdata, fullAnnotatedRegions.Rdata,
## fullCoveragePrep.Rdata with mergePrep).  The per-chromosome .Rdata files are read as the reference reads them; the
## two rankings -- .calculateFWER (:380-386) and .calcPval on the pooled nulls (:302-307) -- run in the library.  When the
## chromosomes were analysed in THIS R session their device handles are still attached to the results
## (attr "dfb_regions"/"dfb_nulls") and `chrResults` can be passed instead of reading files: the pooling is then ONE
## library call over every GPU of the communicator (.mergeResultsGPU), the null lists never leave their GPU.
mergeResults <- function(chrs = c(seq_len(22), "X", "Y"), prefix = ".", significantCut = c(0.05, 0.10), genomicState,
    minoverlap = 20, mergePrep = FALSE, ...) {
    stopifnot(length(significantCut) == 2 & all(significantCut >= 0 & significantCut <= 1))              # :112-114
    verbose <- derfinder:::.advanced_argument("verbose", TRUE, ...)
    optionsStats <- derfinder:::.advanced_argument("optionsStats", NULL, ...)
    cutoffFstatUsed <- derfinder:::.advanced_argument("cutoffFstatUsed", optionsStats$cutoffFstatUsed, ...)
    chrResults <- derfinder:::.advanced_argument("chrResults", NULL, ...)
    chrs <- derfinder::extendedMapSeqlevels(chrs, ...)                                                   # :140 (UCSC names)
    if (!is.null(chrResults) && is.null(names(chrResults))) names(chrResults) <- chrs
    optionsMerge <- list(chrs = chrs, significantCut = significantCut, minoverlap = minoverlap, mergeCall = match.call(),
        cutoffFstatUsed = cutoffFstatUsed, ...)                                                          # :143-147
    save(optionsMerge, file = file.path(prefix, "optionsMerge.Rdata"))
    fullCoveragePrep <- fullTime <- fullNullPermutation <- fullNullWidths <- fullNullStats <- fullFstats <- fullAnno <-
        fullRegs <- structure(vector("list", length(chrs)), names = chrs)
    for (chr in chrs) {                                                                                  # :161-191
        e <- new.env()
        load(file.path(prefix, chr, "fstats.Rdata"), envir = e); fullFstats[[chr]] <- e$fstats
        if (is.null(chrResults)) load(file.path(prefix, chr, "regions.Rdata"), envir = e) else e$regions <- chrResults[[chr]]
        fullRegs[[chr]] <- e$regions$regions; fullNullStats[[chr]] <- e$regions$nullStats
        fullNullWidths[[chr]] <- e$regions$nullWidths; fullNullPermutation[[chr]] <- e$regions$nullPermutation
        load(file.path(prefix, chr, "annotation.Rdata"), envir = e); fullAnno[[chr]] <- e$annotation
        load(file.path(prefix, chr, "timeinfo.Rdata"), envir = e); fullTime[[chr]] <- e$timeinfo
        if (mergePrep) { load(file.path(prefix, chr, "coveragePrep.Rdata"), envir = e); fullCoveragePrep[[chr]] <- e$prep }
    }
    fullRegions <- unlist(GenomicRanges::GRangesList(fullRegs[!vapply(fullRegs, is.null, logical(1))]), use.names = FALSE)
    fullAnnotation <- do.call(rbind, fullAnno)                                                           # :202-214
    if (!is.null(fullAnnotation)) {
        colnames(fullAnnotation)[which(colnames(fullAnnotation) == "strand")] <- "annoStrand"
        rownames(fullAnnotation) <- NULL
        fullAnnotation$name <- as.character(fullAnnotation$name)
        fullAnnotation$annotation <- as.character(fullAnnotation$annotation)
        S4Vectors::values(fullRegions) <- cbind(S4Vectors::values(fullRegions), S4Vectors::DataFrame(fullAnnotation, check.names = FALSE))
    }
    nulls <- unlist(IRanges::RleList(fullNullStats), use.names = FALSE)                                  # :218-238
    widths <- unlist(IRanges::RleList(fullNullWidths), use.names = FALSE)
    permutations <- unlist(IRanges::RleList(fullNullPermutation), use.names = FALSE)
    howMany <- unlist(lapply(fullNullStats, length))
    if (length(nulls) > 0) {
        fullNullSummary <- S4Vectors::DataFrame(stat = nulls, width = widths, chr = S4Vectors::Rle(names(fullNullStats), howMany),
            permutation = permutations, check.names = FALSE)
        fullNullSummary$area <- abs(fullNullSummary$stat) * fullNullSummary$width
        fullNullSummary <- fullNullSummary[order(fullNullSummary$area, decreasing = TRUE), ]
    } else fullNullSummary <- S4Vectors::DataFrame(NULL)
    if (is.null(cutoffFstatUsed) && !is.null(optionsStats)) {                                            # :241-263
        stopifnot(all(c("cutoffType", "cutoffFstat", "models") %in% names(optionsStats)))
        cutoffFstatUsed <- switch(optionsStats$cutoffType,
            empirical = .Call(dfbR_quantile, .engine(), as.numeric(unlist(fullFstats)), as.numeric(optionsStats$cutoffFstat)),
            theoretical = { m <- optionsStats$models; qf(optionsStats$cutoffFstat, ncol(m$mod) - ncol(m$mod0),
                                                         nrow(m$mod) - ncol(m$mod), lower.tail = FALSE) },
            manual = optionsStats$cutoffFstat)
    }
    pooled <- NULL
    if (!is.null(chrResults) && !is.null(cutoffFstatUsed) && nrow(fullNullSummary) > 0) {
        ## handles alive: one collective library call, p-values and FWER per chromosome in calculatePvalues() order
        stopifnot("nPermute" %in% names(optionsStats))
        pooled <- .mergeResultsGPU(chrResults[chrs], optionsStats$nPermute, cutoffFstatUsed, mergeArea = TRUE)
    }
    if (!is.null(cutoffFstatUsed)) {                                                                     # :266-289
        if (nrow(fullNullSummary) > 0) {
            stopifnot(!is.null(optionsStats)); stopifnot("nPermute" %in% names(optionsStats))
            fullRegions$fwer <- if (!is.null(pooled)) unlist(pooled$fwer, use.names = FALSE) else
                .Call(dfbR_calc_fwer, .engine(), as.numeric(cutoffFstatUsed), as.numeric(fullNullSummary$area),
                      as.integer(fullNullSummary$permutation), as.integer(optionsStats$nPermute), as.numeric(fullRegions$area))
            fullRegions$significantFWER <- factor(fullRegions$fwer < significantCut[1], levels = c(TRUE, FALSE))
        } else if (verbose) warning("No null regions were found, so the FWER calculation step will be skipped.")
    }
    save(fullNullSummary, file = file.path(prefix, "fullNullSummary.Rdata"))
    if (!is.null(pooled)) fullRegions$pvalues <- unlist(pooled$pvalues, use.names = FALSE)
    fullRegions <- fullRegions[order(fullRegions$area, decreasing = TRUE)]                               # :300
    if (nrow(fullNullSummary) > 0) {                                                                     # :302-327
        if (is.null(pooled))
            fullRegions$pvalues <- .Call(dfbR_calc_pval, .engine(), as.numeric(fullRegions$area), as.numeric(fullNullSummary$area))
        fullRegions$significant <- factor(fullRegions$pvalues < significantCut[1], levels = c(TRUE, FALSE))
        qvalues <- tryCatch(qvalue::qvalue(fullRegions$pvalues), error = function(e) NULL)
        if (!is.null(qvalues)) {
            qvalues <- qvalues$qvalues; sigQval <- factor(qvalues < significantCut[2], levels = c(TRUE, FALSE))
        } else {
            qvalues <- rep(NA, length(fullRegions$pvalues)); sigQval <- rep(NA, length(fullRegions$pvalues))
        }
        fullRegions$qvalues <- qvalues; fullRegions$significantQval <- sigQval
    }
    save(fullRegions, file = file.path(prefix, "fullRegions.Rdata"))
    fullAnnotatedRegions <- derfinder::annotateRegions(regions = fullRegions, genomicState = genomicState,
        minoverlap = minoverlap, annotate = TRUE, ...)                                                   # :337-343
    save(fullAnnotatedRegions, file = file.path(prefix, "fullAnnotatedRegions.Rdata"))
    save(fullFstats, file = file.path(prefix, "fullFstats.Rdata"))
    save(fullTime, file = file.path(prefix, "fullTime.Rdata"))
    if (mergePrep) save(fullCoveragePrep, file = file.path(prefix, "fullCoveragePrep.Rdata"))
    invisible(prefix)
}
