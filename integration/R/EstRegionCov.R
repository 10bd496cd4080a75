## Drop-in body for R/EstRegionCov.R of the reference: same signature and return list.  What ran
## over the dense gene x cell matrix there runs on the device here, on the dgCMatrix slots:
##   apply(mtx,1,var) / rowSums (:42-43)            -> cs_R_gene_moments  (per-gene sum, sum of squares)
##   colSums(rna[gene_ind,]) per segment (:88-93)   -> cs_R_bin_counts    (gene -> segment map)
##   colSums(rna_control) library sizes (:131-137)  -> cs_R_bin_counts    (every control gene -> one bin)
## Overlaps, quantiles, medians, ratios, names and the optional histograms stay in R.
EstRegionCov=function(mtx=NULL, barcodes=NULL, features=NULL, bed=NULL, celltype0=NULL, var_pt=0.99, var_pt_ctrl=0.99,ngene_filter=0, include='tumor',
                      alpha_source='all', ctrl_region=NULL, seg_table_filtered=NULL,size=NULL, plot_path=NULL, breaks=30){
  if(!include %in% c('tumor','control','all')) stop("include variable is not valid!")
  if(!alpha_source %in% c('all','control')) stop("alpha_source is not valid!")
  if(!grepl("chr",bed[1,1])) bed[,1]=paste0('chr', bed[,1])
  if(is.null(celltype0)) celltype0=cbind(barcodes[,1], rep("normal", nrow(barcodes)))
  M=as(mtx, "CsparseMatrix")
  M=M[,match(celltype0[,1], barcodes[,1]), drop=FALSE]                                     # :36-40
  G=nrow(M); N=ncol(M); cells=celltype0[,1]
  mom=.Call("cs_R_gene_moments", M@i, as.numeric(M@x), as.integer(G), PACKAGE="Clonalscope")
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
  seg=.Call("cs_R_bin_counts", M@p, M@i, as.numeric(M@x), as.integer(G), as.integer(c(0L,cumsum(tabulate(ov[,2], nbins=G)))),
            as.integer(ov[,1]-1L), as.integer(S), PACKAGE="Clonalscope")
  segsum=as.matrix(new("dgCMatrix", p=seg[[1]], i=seg[[2]], x=seg[[3]], Dim=c(S, N)))     # S x N, small
  libg=if(alpha_source=='all') keep_c else keep[which(bed_sub[,1] %in% ctrl_region)]
  flag=integer(G); flag[libg]=1L
  lib=.Call("cs_R_bin_counts", M@p, M@i, as.numeric(M@x), as.integer(G), as.integer(c(0L,cumsum(flag))),
            integer(sum(flag)), 1L, PACKAGE="Clonalscope")
  alpha_all=numeric(N); alpha_all[which(diff(lib[[1]])>0)]=lib[[3]]
  deltas_all=list(); ngenes=c(); sel_ind=c(); alphas=NULL; barcodes_normal=NULL
  if(!is.null(plot_path)){ pdf(plot_path, width = 8, height = 6); par(mfrow=c(2,3)) }
  for(si in seq_len(S)){
    chrr=regs[si]
    sel_cell=which(segsum[si,]>0)                                                          # :93
    if(length(sel_cell)==0) next
    sel_ind=c(sel_ind, chrr)
    types=celltype0[sel_cell,2]
    normal=which(types=='normal')
    tsel=switch(include, tumor=which(types=='tumor'), control=normal, all=seq_along(sel_cell))
    barcodes_normal=cells[sel_cell][normal]
    alphas=alpha_all[sel_cell]; names(alphas)=cells[sel_cell]                              # :131-137
    alpha_controls=alphas[normal]/median(alphas); alpha_tests=alphas[tsel]/median(alphas)  # :141-148
    deltas=(segsum[si,sel_cell][tsel]/alpha_tests)/median(segsum[si,sel_cell][normal]/alpha_controls)  # :153-159
    names(deltas)=cells[sel_cell][tsel]
    if(!is.null(plot_path)){
      tp=types[tsel]
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
