## Drop-in body for R/BayesNonparAlleleCluster.R of the reference: same signature and return
## object; the genotype grid and the priors (:26-98) stay in R, the Z0 default (:99-105), L0 (:131)
## and the sampler (:133-290, with genotype_neighbor) are .Calls into libclonalscope_b200.so.
BayesNonparAlleleCluster=function(Xir_cov=NULL,Xir_allele=NULL, cna_states_WGS=NULL,alpha=1, beta=1, niter=200, sigmas0_cov=NULL, sigmas0_allele=NULL, U0=NULL, Z0=NULL, maxcp=6,seed=200,
                                  rng=getOption("clonalscope.rng", "R"), device=getOption("clonalscope.device", -1L)){
  mu=do.call(rbind, lapply(1:maxcp, function(ii) cbind(rep(0.5*ii,ii+1), c(0,(1:ii)/ii))))  # :26-30 (copies, allele fraction)
  N=nrow(Xir_cov); R=ncol(Xir_cov)
  Xir_cov=as.matrix(Xir_cov); Xir_cov0=Xir_cov
  Xir_cov=apply(Xir_cov, 2, function(x) pmin(x, 3)); storage.mode(Xir_cov)="double"
  Xir_allele=as.matrix(Xir_allele); storage.mode(Xir_allele)="double"
  cna_states_WGS=as.numeric(cna_states_WGS)
  if(is.null(U0)){
    wU0=FALSE
    if(length(cna_states_WGS)>0) U0=matrix(c(rep(4,R), cna_states_WGS), byrow=T, ncol=R)
    else{ U0=matrix(rep(4,R), byrow=T, ncol=R); cna_states_WGS=U0[1,] }
  }else if(ncol(U0)==R){
    wU0=TRUE
    U0=U0[1:max(which(!is.na(U0[,1]))),, drop=F]
    if(nrow(U0)!=1) cna_states_WGS=U0[which(!is.na(result$Uest[,1]))[2],] else cna_states_WGS=U0[1,]   # (the reference's own undefined `result`, :62)
  }else stop("Please specify a U0 matrix with ncol the same as that of Xir!")
  fill=function(s, d, what) if(is.null(s)) rep(d,R) else if(length(s)==1) rep(s,R) else if(length(s)==R) s else stop(paste0("Please specify ",what," with length R or length 1."))
  sigmas0_cov=fill(sigmas0_cov, 0.3, "sigmas0_cov"); sigmas0_allele=fill(sigmas0_allele, 0.15, "sigmas0_allele")
  if(is.null(Z0)){
    PP=.Call("cs_R_loglik_matrix", Xir_cov, matrix(mu[U0,1], ncol=R), as.numeric(sigmas0_cov), PACKAGE="Clonalscope")+
       .Call("cs_R_loglik_matrix", Xir_allele, matrix(mu[U0,2], ncol=R), as.numeric(sigmas0_allele), PACKAGE="Clonalscope")
    Z0=apply(PP, 1, which.max)
  }else if(length(Z0)!=N) stop("Please specify a vector with length the same as the number of cells!")
  else Z0[is.na(Z0)]=0
  colnames(U0)=names(sigmas0_cov)=names(sigmas0_allele)=paste0("R", 1:R)
  rownames(U0)=paste0("Cluster", 1:nrow(U0))
  set.seed(seed)
  U0i=U0; storage.mode(U0i)="integer"
  res=.Call("cs_R_ncrp_gibbs_allele", Xir_cov, Xir_allele, U0i, as.integer(Z0), as.numeric(sigmas0_cov), as.numeric(sigmas0_allele),
            as.numeric(alpha), as.numeric(beta), as.integer(niter), as.integer(seed), as.integer(maxcp),
            as.integer(c(R=0L, philox=1L)[[rng]]), as.integer(device), PACKAGE="Clonalscope")
  unpack=function(flat, tt){                       # rows of sweep tt -> kt x R matrix, emptied clusters all zero (:203-205)
    idx=seq_len(res$u_off[tt+1]-res$u_off[tt])+res$u_off[tt]
    Ut=matrix(0, nrow=res$kt[tt], ncol=R)
    Ut[res$u_labels[idx],]=matrix(flat, ncol=R, byrow=TRUE)[idx,,drop=F]
    colnames(Ut)=paste0("R", 1:R); rownames(Ut)=paste0('Cluster',1:nrow(Ut)); Ut
  }
  nm=function(v){ names(v)=paste0("R", 1:R); v }
  Uall=lapply(1:niter, function(tt) unpack(res$u_state, tt))
  Ucov=lapply(1:niter, function(tt) unpack(res$u_cov, tt))
  Uallele=lapply(1:niter, function(tt) unpack(res$u_allele, tt))
  sigma_cov_all=lapply(1:niter, function(tt) nm(res$sigma_cov_all[,tt]))
  sigma_allele_all=lapply(1:niter, function(tt) nm(res$sigma_allele_all[,tt]))
  for(tt in 1:niter) if(tt%%20==0) print(paste0("Iteration:", tt))
  Zall=res$Zall; rownames(Zall)=paste0("Iter",1:niter); colnames(Zall)=rownames(Xir_cov)
  results=list(Zall=Zall, Uall=Uall,Ucov=Ucov,Uallele=Uallele, sigma_cov_all=sigma_cov_all,sigma_allele_all=sigma_allele_all, Likelihood=res$Likelihood)
  priors=list(Z0=Z0, U0=U0, sigmas0_cov=sigmas0_cov, sigmas0_allele=sigmas0_allele, alpha=alpha, beta=beta, cna_states_WGS=cna_states_WGS, L0=res$L0, wU0=wU0)
  return(list(results=results, priors=priors, data=list(Xir_cov=Xir_cov0, Xir_allele=Xir_allele)))
}
