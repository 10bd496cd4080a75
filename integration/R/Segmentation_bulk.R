## Drop-in for the HMM loop of R/Segmentation_bulk.R:150-166 and the pooling loop of
## R/CreateSegtableNoWGS.R:62-69,77.  Everything else in those two functions (name parsing, quantile /
## median clipping, the seg-table rows, the breakpoint union) is host logic that stays as it is.
##
## (1) Segmentation_bulk: replace, inside `for(ii in ...)`,
##         cov5=caTools::runmean(cov4, nmean); z <- HiddenMarkov::dthmm(cov5, Pi, delta, "norm", ...); results <- HiddenMarkov::Viterbi(z)
##     by one call for ALL chromosomes made before the loop:
cs_segment_all=function(cov3, chromnum, chroms, adj, nmean, hmm_states, hmm_sd, hmm_p){
  med=median(cov3)
  seqs=lapply(chroms, function(ii) (cov3/med)[which(chromnum==paste0(ii))]+adj)           # :151
  off=c(0, cumsum(sapply(seqs, length)))
  t=hmm_p
  Pi=matrix(c(1-3*t,t,t,t, t,1-3*t,t,t, t,t,1-3*t,t, t,t,t,1-3*t), byrow=TRUE, nrow=4)      # :160
  res=.Call("cs_R_segment_hmm", as.numeric(off), as.numeric(unlist(seqs)), as.integer(nmean),
            c(hmm_states[3], hmm_states[2], 1, hmm_states[1]), rep(hmm_sd, 4), c(0.1,0.2,0.5,0.2),
            as.numeric(t(Pi)), PACKAGE="Clonalscope")
  ## results for chromosome k: res$state[(off[k]+1):off[k+1]]  (1-based states, like Viterbi())
  lapply(seq_along(chroms), function(k) list(cov4=seqs[[k]], results=res$state[(off[k]+1):off[k+1]]))
}

## (2) CreateSegtableNoWGS: replace the loop
##         for(ii in 1:length(table(Zest))){ sub=bin_mtx[,which(Zest==names(table(Zest))[ii])]; bin_by_cluster[,ii]=rowSums(sub) }
##     and  ref_counts=rowSums(bin_mtx[,which(celltype0[,2]=='normal')])  by
cs_bin_by_cluster=function(bin_mtx, Zest, celltype0){
  M=as(bin_mtx, "CsparseMatrix")
  lev=names(table(Zest))
  label=match(as.character(Zest), lev)-1L; label[is.na(label)]=-1L
  out=.Call("cs_R_rowsum_by_label", M@p, M@i, as.numeric(M@x), nrow(M), as.integer(label), length(lev), PACKAGE="Clonalscope")
  colnames(out)=paste0('C', lev)
  ref=.Call("cs_R_rowsum_by_label", M@p, M@i, as.numeric(M@x), nrow(M), ifelse(celltype0[,2]=='normal', 0L, -1L), 1L, PACKAGE="Clonalscope")
  list(bin_by_cluster=out, ref_counts=ref)
}

## (3) the head of the tutorials (samples/P5931/scRNA/README.md:72): mtx <- readMM("matrix.mtx.gz")
readMM_cs=function(path){
  r=.Call("cs_R_read_mm", path.expand(path), PACKAGE="Clonalscope")
  new("dgCMatrix", Dim=r$Dim, p=r$p, i=r$i, x=r$x)
}
