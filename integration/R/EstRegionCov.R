## Drop-in body for R/EstRegionCov.R of the reference: same signature and return list.  The matrix is
## uploaded once and everything that ran over the dense gene x cell matrix runs on that device copy:
##   mtx[, match(celltype0[,1], barcodes[,1])] (:40)          -> cs_R_csc_upload      (column gather on the device)
##   apply(mtx,1,var) / rowSums (:42-43)                       -> cs_R_csc_gene_moments
##   per segment: colSums, sel_cell, library sizes, both medians, ratios (:88-159), all segments at once
##                                                             -> cs_R_csc_region_cov
## Overlaps, the quantile of the variance filter, names and the optional histograms stay in R.
EstRegionCov=function(mtx=NULL, barcodes=NULL, features=NULL, bed=NULL, celltype0=NULL, var_pt=0.99, var_pt_ctrl=0.99,ngene_filter=0, include='tumor',
                      alpha_source='all', ctrl_region=NULL, seg_table_filtered=NULL,size=NULL, plot_path=NULL, breaks=30){
  if(!include %in% c('tumor','control','all')) stop("include variable is not valid!")
  if(!alpha_source %in% c('all','control')) stop("alpha_source is not valid!")
  if(!grepl("chr",bed[1,1])) bed[,1]=paste0('chr', bed[,1])
  if(is.null(celltype0)) celltype0=cbind(barcodes[,1], rep("normal", nrow(barcodes)))
  M=as(mtx, "CsparseMatrix")
  G=nrow(M); cells=celltype0[,1]; N=length(cells)
  h=.Call("cs_R_csc_upload", M@p, M@i, as.numeric(M@x), as.integer(G), as.integer(match(celltype0[,1], barcodes[,1])-1L), PACKAGE="Clonalscope")  # :36-40
  mom=.Call("cs_R_csc_gene_moments", h, as.integer(G), PACKAGE="Clonalscope")
  ngenes_g=mom[[1]]; rna_var=pmax(mom[[2]]-mom[[1]]^2/N, 0)/(N-1)                          # :42-43
  keep  =which(rna_var<quantile(rna_var,var_pt)      & (rna_var!=0) & ngenes_g>ngene_filter)  # :47
  keep_c=which(rna_var<quantile(rna_var,var_pt_ctrl) & (rna_var!=0) & ngenes_g>ngene_filter)  # :48
  bed_sub=bed[match(features[keep,1], bed[,4]),, drop=F]; bed_sub[is.na(bed_sub)]=0        # :51-52
  est_name=paste0('pt', var_pt, "_",include,"_alpha",alpha_source)
  if(is.null(seg_table_filtered)){
    message("Estimation for each chromosome")
    seg_table_filtered=data.frame("chr"=gsub("chr","",size[1:22,1]), 'start'=0, 'end'=as.numeric(size[1:22,2]),
                                  'states'=0, 'length'=as.numeric(size[1:22,2]),'mean'=0, 'var'=0, 'Var1'=1:22,'Freq'=50000,'chrr'=size[1:22,1], stringsAsFactors = F)
  }
  ## one query per distinct chrr (:80-87); gene -> segment CSR map over the kept genes
  regs=unique(paste0(seg_table_filtered$chrr)); S=length(regs)
  query=GRanges(seqnames=paste0('chr',seg_table_filtered$chr),
                ranges=IRanges(as.numeric(seg_table_filtered$start), as.numeric(seg_table_filtered$end)))
  subject=GRanges(seqnames=bed_sub[,1], ranges=IRanges(start=as.numeric(bed_sub[,2]), end=as.numeric(bed_sub[,3])))
  ov=as.matrix(findOverlaps(query=query, subject=subject))
  ov=cbind(match(paste0(seg_table_filtered$chrr)[ov[,1]], regs), keep[ov[,2]])             # (segment, gene row)
  ov=ov[order(ov[,2], ov[,1]),,drop=FALSE]
  seg_ngene=tabulate(ov[,1], nbins=S)
  libg=if(alpha_source=='all') keep_c else keep[which(bed_sub[,1] %in% ctrl_region)]
  flag=raw(G); flag[libg]=as.raw(1L)
  code=ifelse(celltype0[,2]=='normal', 0L, ifelse(celltype0[,2]=='tumor', 1L, 2L))
  rc=.Call("cs_R_csc_region_cov", h, as.integer(N), as.integer(S), as.integer(c(0L,cumsum(tabulate(ov[,2], nbins=G)))),
           as.integer(ov[,1]-1L), flag, as.integer(code), match(include, c('tumor','control','all'))-1L, PACKAGE="Clonalscope")
  out_type=switch(include, tumor=celltype0[,2]=='tumor', control=celltype0[,2]=='normal', all=rep(TRUE, N))
  deltas_all=list(); ngenes=c(); sel_ind=c(); alphas=NULL; barcodes_normal=NULL
  if(!is.null(plot_path)){ pdf(plot_path, width = 8, height = 6); par(mfrow=c(2,3)) }
  for(si in seq_len(S)){
    chrr=regs[si]
    if(rc$nsel[si]==0) next                                                                # :93 no cell with a positive sum
    sel_ind=c(sel_ind, chrr)
    sel=rc$sums[,si]>0
    tsel=which(sel & out_type)
    types=celltype0[sel,2]
    barcodes_normal=cells[sel & celltype0[,2]=='normal']
    alphas=rc$alpha[sel]; names(alphas)=cells[sel]                                         # :131-137
    deltas=rc$deltas[tsel,si]; names(deltas)=cells[tsel]                                   # :141-159 (device)
    if(!is.null(plot_path)){
      tp=celltype0[tsel,2]
      if(sum(tp=='tumor')>1 & sum(tp=='normal')>1){
        plot(hist(deltas[tp=='tumor'],breaks=breaks,plot=FALSE), xlim=c(0,3), col=rgb(239,90,155,max=255,alpha=50),
             main=paste0(chrr, ":", seg_ngene[si], " genes; ", length(deltas), " cells"))
        plot(hist(deltas[tp=='normal'],breaks=breaks,plot=FALSE), add=TRUE, col=rgb(56,101,195,max=255,alpha=50))
      }else if(length(deltas)>1) plot(hist(deltas,breaks=breaks,plot=FALSE), xlim=c(0,3), main=paste0(chrr, ":", seg_ngene[si], " genes; ", length(deltas)))
    }
    deltas_all[[paste0(est_name,chrr)]]=deltas
    ngene=seg_ngene[si]; names(ngene)=paste0(est_name,chrr); ngenes=c(ngenes, ngene)
    print(chrr)
  }
  if(!is.null(plot_path)) dev.off()
  seg_table_filtered=seg_table_filtered[which(seg_table_filtered$chrr %in% sel_ind),, drop=F]
  return(list(deltas_all=deltas_all, ngenes=ngenes, seg_table_filtered=seg_table_filtered, alpha=alphas, barcodes_normal=barcodes_normal))
}
