## Drop-in body for R/Gene_bin_cell_rna_filtered.R of the reference: same signature, same
## return object (bin x cell dgCMatrix with the reference's dimnames).  The interval logic
## (:28-64) stays R's own (GenomicRanges); the dense per-bin loop (:66-81) is one .Call on the
## dgCMatrix slots.  The gene x cell matrix is never densified.
Gen_bin_cell_rna_filtered=function(bin_bed=NULL, barcodes=NULL, gene_matrix=NULL, genes=NULL, gtf=NULL){
  barcodes=barcodes[,1]
  if(typeof(genes)!="list") genes=as.data.frame(genes)
  cat("Read gene by cell matrix...\n")
  cat(paste0("Total ", length(barcodes)," cells.\n"))
  cat(paste0("Total ", nrow(genes)," genes.\n"))
  if(grepl("chr",bin_bed[1,1])) bin_bed[,1]=gsub("chr","",bin_bed[,1])                    # :28-33
  tiles=bin_bed[order(as.numeric(bin_bed[,1]), as.numeric(bin_bed[,2])),]                 # :35-36
  query=GRanges(paste0('chr',tiles[,1]), IRanges(tiles[,2]+1,tiles[,3]))
  if(is.null(gtf$gene_id)){                                                               # :42-49
    gtf=gtf[which(gtf$V3=='gene'),]
    ensg=sapply(strsplit(sapply(strsplit(gtf$V9,';'),'[',1),'gene_id '),'[',2)
  }else ensg=gtf$gene_id
  span=paste0(gtf$V1,'-',gtf$V4,'-', gtf$V5); names(span)=ensg
  genes$site=span[genes[,1]]
  genes[is.na(genes$site),"site"]="0-0-0"
  parts=strsplit(genes$site,'-')
  sq=sapply(parts,'[',1); if(!grepl("chr",genes$site[1])) sq=paste0('chr',sq)
  subject=GRanges(sq, IRanges(as.numeric(sapply(parts,'[',2)), as.numeric(sapply(parts,'[',3))))
  ov=as.matrix(findOverlaps(query, subject))                                              # :56-57
  cat("Generate bin-by-cell matrix...")
  ## gene -> bins as a CSR map (0-based bins, grouped by gene, bins ascending within a gene)
  G=nrow(genes); B=nrow(tiles)
  ov=ov[order(ov[,2], ov[,1]),,drop=FALSE]
  map_ptr=c(0L, cumsum(tabulate(ov[,2], nbins=G)))
  M=as(gene_matrix, "CsparseMatrix"); if(!is(M,"dgCMatrix")) M=as(M,"dMatrix")
  res=.Call("cs_R_bin_counts", M@p, M@i, as.numeric(M@x), as.integer(G), as.integer(map_ptr),
            as.integer(ov[,1]-1L), as.integer(B), PACKAGE="Clonalscope")
  chr200_mtx=new("dgCMatrix", p=res[[1]], i=res[[2]], x=res[[3]], Dim=c(B, ncol(M)),
                 Dimnames=list(paste0("chr",tiles[,1],"-", tiles[,2],"-", tiles[,3]), barcodes)) # :79-81
  message("The bin-by-cell matrix has beed successfully generated!")
  return(chr200_mtx)
}
