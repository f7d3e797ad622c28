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
    mapped <- if (!is.null(totalMapped)) { stopifnot(length(totalMapped) == ncol(data)); totalMapped / targetSize }
    r <- .rle_lists(data)
    prev <- if (!is.null(index)) .lgl_runs(index)
    kind <- if (is.null(cutoff)) 0L else if (filter == "one") 1L else 2L
    h <- .Call(dfbR_filter, .engine(), r$values, r$lengths, nrow(data), mapped, kind,
               if (is.null(cutoff)) 0 else as.numeric(cutoff), prev$lengths, prev$first)
    res <- list(coverage = if (returnCoverage) .cov_dataframe(h, names(data)) else NULL,
                position = .as_lgl_rle(.Call(dfbR_cov_position, h)))
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
    if (!is.null(colsubset)) {                                                             # :145-160
        sub <- filterData(coverageInfo$coverage[, colsubset, drop = FALSE], cutoff = cutoff,
                          index = coverageInfo$position, ...)
        h <- attr(sub, "dfb_cov"); coverageInfo <- sub
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
        meanCoverage = S4Vectors::Rle(.Call(dfbR_cov_vector, h, 2L, 0L)),
        groupMeans = if (!is.null(groupInfo)) structure(lapply(seq_len(nlevels(groupInfo)) - 1L, function(g)
            S4Vectors::Rle(.Call(dfbR_cov_vector, h, 3L, g))), names = levels(groupInfo)) else list())
    attr(res, "dfb_cov") <- h
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
    if (smooth) stop("smooth = TRUE is not supported by the B200 engine")
    basic <- derfinder:::.advanced_argument("basic", FALSE, ...)
    maxRegionGap <- derfinder:::.advanced_argument("maxRegionGap", 0L, ...)
    if (!basic) stopifnot(!is.null(position))                                            # :135-137
    if (maxClusterGap < maxRegionGap)
        warning("'maxClusterGap' is less than 'maxRegionGap' which nullifies it's intended use.")
    cp <- .cutoff_pair(as.numeric(cutoff)); p <- .lgl_runs(position)
    x <- .Call(dfbR_find_regions, .engine(), as.numeric(fstats), p$lengths, p$first, cp[1], cp[2],
               as.integer(maxRegionGap), as.integer(maxClusterGap), basic)
    if (is.null(x)) { warning("Found no regions. Consider changing the F-statistics cutoff, or increasing 'maxRegionGap' if you are smoothing."); return(NULL) }
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
    x <- .Call(dfbR_calculate_pvalues, .engine(), h, if (cached) NULL else as.numeric(fstats), models$mod,
               models$mod0, adjustF, perm, cp[1], cp[2], as.integer(maxRegionGap), as.integer(maxClusterGap),
               length(coveragePrep$groupMeans))
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
    list(regions = regs, nullStats = S4Vectors::Rle(x$nullStats), nullWidths = S4Vectors::Rle(x$nullWidths),
         nullPermutation = S4Vectors::Rle(x$nullPermutation))                            # :421-424
}

## ---- regionMatrix per chromosome (R/regionMatrix.R:170-283), coverageMatrix only
.regionMatrixByChr <- function(covInfo, chr, cutoff, L, totalMapped, targetSize, maxRegionGap = 0L, maxClusterGap = 300L) {
    filt <- filterData(covInfo$coverage, cutoff = cutoff, filter = "mean", totalMapped = totalMapped,
                       targetSize = targetSize, returnMean = TRUE, returnCoverage = FALSE)
    h <- attr(filt, "dfb_cov")
    regs <- findRegions(position = filt$position, fstats = filt$meanCoverage, chr = chr, cutoff = cutoff,
                        maxRegionGap = maxRegionGap, maxClusterGap = maxClusterGap)
    if (is.null(regs)) return(list(regions = GenomicRanges::GRanges(), coverageMatrix = NULL))             # :230-239
    m <- .Call(dfbR_region_matrix, .engine(), h, regs$indexStart, regs$indexEnd, as.numeric(L))
    colnames(m) <- names(covInfo$coverage)
    list(regions = regs, coverageMatrix = m)
}

## ---- regionMatrix (R/regionMatrix.R:124-168): chromosomes one after the other on the GPU (the reference's bpmapply
## over chromosomes :164 would fork the CUDA context); returnBP = TRUE is served by the reference's own
## getRegionCoverage() on the returned regions
regionMatrix <- function(fullCov, cutoff = 5, L, totalMapped = 80e6, targetSize = 80e6, runFilter = TRUE,
    returnBP = TRUE, ...) {
    stopifnot(!is.null(cutoff), !is.null(names(fullCov)))                                 # :128-131
    if (!runFilter) stop("runFilter = FALSE: pass the unfiltered coverage; the filter runs on the device")
    maxRegionGap <- derfinder:::.advanced_argument("maxRegionGap", 0L, ...)
    maxClusterGap <- derfinder:::.advanced_argument("maxClusterGap", 300L, ...)
    res <- lapply(names(fullCov), function(chr) {
        covInfo <- fullCov[[chr]]
        if (!is.list(covInfo) || is(covInfo, "DataFrame")) covInfo <- list(coverage = covInfo)
        out <- .regionMatrixByChr(covInfo, chr, cutoff, L, totalMapped, targetSize, maxRegionGap, maxClusterGap)
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

.calcPval <- function(areas, nullareas) .Call(dfbR_calc_pval, .engine(), as.numeric(areas), as.numeric(nullareas))
.calculateFWER <- function(cutoffFstatUsed, areaNull, permutation, nPermute, areaReg)
    .Call(dfbR_calc_fwer, .engine(), cutoffFstatUsed, as.numeric(areaNull), as.integer(permutation),
          as.integer(nPermute), as.numeric(areaReg))
