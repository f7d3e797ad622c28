library("testthat")
library("derfinderB200")
test_check("derfinderB200")
