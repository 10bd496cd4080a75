## Drop-in body for the coverage branch of R/AssignCluster.R of the reference (:22-208): same
## signature and return list.  The tally (:37), the cluster means (:46-51,67-72), the sigmas
## (:53,77) and the orphans' log-likelihoods (:58) are .Calls; set.seed(100) and sample.int stay
## R's own, so the re-draw of the orphans consumes R's stream exactly as the reference does.
## allele=TRUE falls through to the reference's own body (AssignCluster_reference, the renamed
## original), which this file does not replace.
AssignCluster=function(cluster_obj=NULL, mincell=20, cutoff=0.2, allele=FALSE, cna_states_WGS=NULL, rm_extreme=F, cap_corr=1.5){
  if(allele!=FALSE) return(AssignCluster_reference(cluster_obj, mincell, cutoff, allele, cna_states_WGS, rm_extreme, cap_corr))
  R=ncol(cluster_obj$priors$U0)
  Zall=cluster_obj$results$Zall; storage.mode(Zall)="integer"
  Uall=cluster_obj$results$Uall
  Xir=as.matrix(cluster_obj$data); storage.mode(Xir)="double"
  N=nrow(Xir); K=nrow(Uall[[length(Uall)]])
  set.seed(100)
  maj_vote=.Call("cs_R_majority_vote", Zall, PACKAGE="Clonalscope"); names(maj_vote)=colnames(Zall)   # :37
  Zest=maj_vote
  ind_clusters=as.numeric(names(table(Zest))[which(table(Zest)>mincell)])                  # :40-43
  ind_cell=which(!Zest %in% ind_clusters)
  Zest[ind_cell]=NA
  means=function(Z){                                                                       # colMeans(..., na.rm=T), NaN -> 0
    st=.Call("cs_R_cluster_stats", Xir, as.integer(Z), as.integer(K), PACKAGE="Clonalscope")
    Usub=matrix(NA_real_, nrow=K, ncol=R)
    m=st$sum[ind_clusters,,drop=F]/st$count[ind_clusters,,drop=F]; m[is.na(m)]=0
    Usub[ind_clusters,]=m; Usub
  }
  sig=function(Z, Usub, denom){ U=Usub; U[is.na(U)]=0
    sqrt(.Call("cs_R_sigma_stats", Xir, as.integer(Z), U, PACKAGE="Clonalscope")$sq/denom) }
  Usub=means(Zest)
  sigmas=sig(Zest, Usub, N-length(ind_cell))                                               # :53
  if(length(ind_cell)>0){                                                                  # :56-64
    PP=.Call("cs_R_loglik_matrix", Xir[ind_cell,,drop=F], Usub[ind_clusters,,drop=F], as.numeric(sigmas), PACKAGE="Clonalscope")
    for(jj in seq_along(ind_cell)){
      P=rep(NA, K); P[ind_clusters]=PP[jj,]-max(PP[jj,], na.rm=T)
      P[is.na(P)]=min(P, na.rm=T)-100
      Zest[ind_cell[jj]]=sample.int(K, 1, replace=F, prob=exp(P))
    }
  }
  Usub=means(Zest); colnames(Usub)=colnames(Xir); rownames(Usub)=paste0("Cluster", 1:K)
  sigmas=sig(Zest, Usub, N); names(sigmas)=colnames(Xir)                                   # :77
  priors=if(is.null(cna_states_WGS)) cluster_obj$priors$cna_states_WGS else cna_states_WGS
  if(all(priors==1)) priors=rep(1.5, R)
  Usub2=pmax(pmin(Usub, cap_corr), 0.5)                                                    # :106-107
  corrs=apply(Usub2, 1, function(x){
    tmp=sort(((as.numeric(x)-1)*(as.numeric(priors)-1))/(sqrt(sum((as.numeric(x)-1)^2))*sqrt(sum((as.numeric(priors)-1)^2))))
    if(length(tmp)==0) NA else if(rm_extreme==T) NULL else sum(tmp) })
  corrs=as.numeric(sapply(corrs, function(v) if(is.null(v)) NA else v)); names(corrs)=rownames(Usub)
  corrs[is.na(corrs)]=0
  if(max(corrs, na.rm=T)<0.5) warning("WGS and scRNA do not match well. annot might not be accurate!")
  annot=as.numeric(corrs)
  if(cutoff=='km'){ km=kmeans(annot[which(annot!=100)],2); cutoff=mean(km$centers) }       # :196-199
  annot[annot>cutoff]="T"; annot[annot!="T"]=paste0("N")
  message("Succeed!")
  return(list(Zest=Zest, corrs=corrs[Zest], annot=annot[Zest],Uest=Usub, sigmas_est=sigmas, Zmaj= maj_vote, cutoff=cutoff, U0=cluster_obj$priors$U0, mincell=mincell, wU0=cluster_obj$priors$wU0))
}
