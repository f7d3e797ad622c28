## Parity of the B200 engine's R glue with the reference package, on the reference's own objects and known
## answers (lcolladotor/derfinder tests/testthat/test_Fstats.R, test_findRegions.R, test_ER_level.R, test_data.R).
## Needs R, the derfinder package (data sets + the host-side helpers that stay in R) and a B200; skipped otherwise.
## Every expectation compares the engine's function with the reference's function of the same name on the same input.

skip_if_not_installed("derfinder")
skip_if(inherits(try(derfinderB200:::.engine(), silent = TRUE), "try-error"), "no B200 device")

library("GenomicRanges")
ref <- asNamespace("derfinder")
genomeData <- derfinder::genomeData; genomeInfo <- derfinder::genomeInfo
genomeFstats <- derfinder::genomeFstats; genomeRegions <- derfinder::genomeRegions

collapsedFull <- derfinder::collapseFullCoverage(list(genomeData$coverage), verbose = FALSE)
sampleDepths <- derfinder::sampleDepth(collapsedFull, probs = 0.5, verbose = FALSE)
models <- derfinder::makeModels(sampleDepths, testvars = genomeInfo$pop, adjustvars = data.frame(genomeInfo$gender))

test_that("filterData matches the reference (test_data.R:174-191)", {
    raw <- derfinder::genomeDataRaw$coverage
    for (args in list(list(cutoff = 0), list(cutoff = 10), list(cutoff = 0.5, filter = "mean"), list(cutoff = NULL),
                      list(cutoff = 5, totalMapped = rep(40e6, 31), targetSize = 80e6),
                      list(cutoff = 5, totalMapped = 80e6, targetSize = 80e6),      # scalar totalMapped is recycled
                      list(cutoff = 5, totalMapped = rep(40e6, 31), targetSize = 0))) {  # targetSize = 0: no normalisation
        a <- do.call(derfinderB200::filterData, c(list(data = raw, returnMean = TRUE), args))
        b <- do.call(derfinder::filterData, c(list(data = raw, returnMean = TRUE, verbose = FALSE), args))
        expect_equivalent(a$coverage, b$coverage); expect_equal(a$position, b$position)
        expect_equal(a$meanCoverage, b$meanCoverage)
    }
    f1 <- derfinderB200::filterData(raw, cutoff = 0)
    expect_equal(derfinderB200::filterData(f1$coverage[, 1:2], 5, index = f1$position)$position,
                 derfinder::filterData(f1$coverage[, 1:2], 5, index = f1$position, verbose = FALSE)$position)
})

prep <- derfinderB200::preprocessCoverage(genomeData, cutoff = 0, scalefac = 32, chunksize = 1e3, colsubset = NULL)
test_that("preprocessCoverage (test_findRegions.R:25-44)", {
    rp <- derfinder::preprocessCoverage(genomeData, cutoff = 0, scalefac = 32, chunksize = 1e3, colsubset = NULL, verbose = FALSE)
    expect_equal(prep$position, rp$position); expect_equivalent(prep$coverageProcessed, rp$coverageProcessed)
    expect_equal(prep$mclapplyIndex, rp$mclapplyIndex); expect_equal(prep$meanCoverage, rp$meanCoverage)
    a <- derfinderB200::preprocessCoverage(genomeData, cutoff = 1, scalefac = 32, chunksize = 1e3, colsubset = 1:2,
                                           groupInfo = factor(letters[1:2]))
    b <- derfinder::preprocessCoverage(genomeData, cutoff = 1, scalefac = 32, chunksize = 1e3, colsubset = 1:2,
                                       groupInfo = factor(letters[1:2]), verbose = FALSE)
    expect_equal(a$position, b$position); expect_equivalent(a$coverageProcessed, b$coverageProcessed)
    expect_equal(a$groupMeans, b$groupMeans); expect_equal(a$meanCoverage, b$meanCoverage)
})

fstats <- derfinderB200::calculateStats(prep, models)
test_that("calculateStats (test_Fstats.R:43-68)", {
    expect_equal(as.numeric(fstats), as.numeric(genomeFstats), tolerance = 1e-6)
    expect_error(derfinderB200::calculateStats(prep, list(mod = models$mod[-1, ], mod0 = models$mod0[-1, ])))
})

test_that("findRegions (test_findRegions.R:49-68, test-maxRegionGap.R)", {
    for (args in list(list(), list(cutoff = 1), list(cutoff = 0, maxRegionGap = 1e5, maxClusterGap = 1e6),
                      list(cutoff = 1, maxRegionGap = 50L), list(cutoff = 1, basic = TRUE))) {
        a <- do.call(derfinderB200::findRegions, c(list(position = prep$position, fstats = genomeFstats, chr = "chr21"), args))
        b <- do.call(derfinder::findRegions, c(list(position = prep$position, fstats = genomeFstats, chr = "chr21", verbose = FALSE), args))
        expect_equivalent(a, b)
    }
    seg <- ref$.clusterMakerRle(prep$position, 0L, ranges = TRUE)
    expect_equivalent(derfinderB200::findRegions(fstats = genomeFstats, chr = "chr21", cutoff = 1, segmentIR = seg, basic = TRUE),
                      derfinder::findRegions(fstats = genomeFstats, chr = "chr21", cutoff = 1, segmentIR = seg, basic = TRUE, verbose = FALSE))
    expect_equal(suppressWarnings(derfinderB200::findRegions(position = Rle(FALSE, 10), fstats = Rle(0, 10), chr = "chr21")), NULL)
})

test_that("calculatePvalues reproduces genomeRegions (test_Fstats.R:80-90)", {
    suppressWarnings(RNGversion("3.5.0"))
    got <- derfinderB200::calculatePvalues(prep, models, genomeFstats, nPermute = 10, seeds = 1:10, chr = "chr21", cutoff = 1)
    expect_equal(ranges(got$regions), ranges(genomeRegions$regions))
    for (k in c("value", "area", "indexStart", "indexEnd", "meanCoverage", "meanCEU", "meanYRI"))
        expect_equal(mcols(got$regions)[[k]], mcols(genomeRegions$regions)[[k]], tolerance = 1e-6)
    expect_equivalent(got$regions$cluster, genomeRegions$regions$cluster)
    expect_equal(as.integer(got$nullWidths), as.integer(genomeRegions$nullWidths))
    expect_equal(as.integer(got$nullPermutation), as.integer(genomeRegions$nullPermutation))
    expect_equal(as.numeric(got$nullStats), as.numeric(genomeRegions$nullStats), tolerance = 1e-6)
    expect_true(sum(got$regions$pvalues == genomeRegions$regions$pvalues) >= 31)   # near-tie band, see DESIGN.md
    expect_equal(derfinderB200:::.calcPval(1:2, 1:4), c(0.8, 0.6))
    expect_equal(derfinderB200:::.calculateFWER(10, 1:4, rep(c(1, 3), each = 2), 3, 1:2), c(1, 0.75))
    m <- derfinderB200:::.mergeResultsGPU(list(chr21 = got), 10, 1, mergeArea = FALSE)   # one chromosome: pooled == own
    expect_equal(m$pvalues$chr21, got$regions$pvalues)
    expect_equal(m$fwer$chr21, ref$.calculateFWER(1, as.numeric(got$nullStats) * as.integer(got$nullWidths),
                                                  as.integer(got$nullPermutation), 10, got$regions$area))
})

test_that("regionMatrix (test_ER_level.R:5-70)", {
    set.seed(20160606)
    x <- Rle(round(runif(1e4, max = 10))); y <- Rle(round(runif(1e4, max = 10))); z <- Rle(round(runif(1e4, max = 10)))
    fullCov <- list("chr21" = DataFrame(x, y, z)); libs <- sapply(fullCov$chr21, sum)
    for (args in list(list(L = 36, totalMapped = libs, targetSize = 4e4), list(L = 36),   # default totalMapped = 80e6 (scalar)
                      list(L = c(36, 50, 76)))) {
        a <- do.call(derfinderB200::regionMatrix, c(list(fullCov = fullCov, maxRegionGap = 10L, maxClusterGap = 300L, returnBP = FALSE), args))
        b <- do.call(derfinder::regionMatrix, c(list(fullCov = fullCov, maxRegionGap = 10L, maxClusterGap = 300L, returnBP = FALSE, verbose = FALSE), args))
        expect_equivalent(a$chr21$regions, b$chr21$regions); expect_equal(a$chr21$coverageMatrix, b$chr21$coverageMatrix)
    }
    filt <- list("chr21" = derfinder::filterData(fullCov$chr21, returnMean = TRUE, filter = "mean", cutoff = 5,
                                                 totalMapped = libs, targetSize = 4e4, verbose = FALSE))
    a <- derfinderB200::regionMatrix(filt, maxRegionGap = 10L, maxClusterGap = 300L, L = 36, runFilter = FALSE, returnBP = FALSE)
    b <- derfinder::regionMatrix(filt, maxRegionGap = 10L, maxClusterGap = 300L, L = 36, runFilter = FALSE, returnBP = FALSE, verbose = FALSE)
    expect_equal(a$chr21$coverageMatrix, b$chr21$coverageMatrix)
})

test_that("mergeResults (test_Fstats.R:94-145): same files, same fullRegions as the reference on the bundled chr21 results", {
    for (d in c("merge-ref", "merge-b200")) {
        dir.create(d, showWarnings = FALSE, recursive = TRUE)
        file.copy(system.file(file.path("extdata", "chr21"), package = "derfinder", mustWork = TRUE), d, recursive = TRUE)
    }
    derfinder::mergeResults(chrs = "21", prefix = "merge-ref", genomicState = derfinder::genomicState$fullGenome, verbose = FALSE)
    derfinderB200::mergeResults(chrs = "21", prefix = "merge-b200", genomicState = derfinder::genomicState$fullGenome, verbose = FALSE)
    expect_equal(sort(dir("merge-b200")), sort(dir("merge-ref")))
    e1 <- new.env(); e2 <- new.env()
    load(file.path("merge-ref", "fullRegions.Rdata"), envir = e1); load(file.path("merge-b200", "fullRegions.Rdata"), envir = e2)
    expect_equivalent(ranges(e2$fullRegions), ranges(e1$fullRegions))
    for (k in c("area", "pvalues", "significant", "qvalues", "significantQval"))
        expect_equal(mcols(e2$fullRegions)[[k]], mcols(e1$fullRegions)[[k]])
    load(file.path("merge-ref", "fullNullSummary.Rdata"), envir = e1); load(file.path("merge-b200", "fullNullSummary.Rdata"), envir = e2)
    expect_equal(as.data.frame(e2$fullNullSummary), as.data.frame(e1$fullNullSummary))
})
